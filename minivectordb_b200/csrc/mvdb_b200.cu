// mvdb_b200.cu -- host side of the C ABI declared in include/mvdb_b200.h.
//
// One mvdb_index = one HBM-resident, growable, row-major fp32 matrix on one
// B200 plus a live-row bitmask (tombstones).  It stands in for the faiss
// IndexFlatIP object and the numpy `embeddings` matrix that the reference's
// VectorDatabase owns (ref minivectordb/vector_database.py:12, 17, 43-46).
//
// Memory: the matrix lives in a virtual-address reservation sized for the
// whole device (cuMemAddressReserve); physical HBM is mapped behind it chunk
// by chunk (cuMemCreate/cuMemMap) as rows arrive.  Appending therefore never
// moves a row and never copies the matrix (the reference pays an O(N*d)
// np.vstack per insert, vector_database.py:72), and a search running
// concurrently with an append keeps valid pointers.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/mvdb_b200.h"
#include "aux_kernels.cuh"
#include "gemm_tc.cuh"
#include "scan.cuh"
#include "scan_i8.cuh"

using namespace mvdb;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local bool tl_allow_pdl = false;    // set by the device-buffer entry points (option "pdl")
static thread_local bool tl_force_scan = false;   // overflow fallback of the batched path: stay on the fp32 scan
static thread_local bool tl_host_checks_i8 = false;   // the caller synchronises and re-runs an overflowed int8 search itself
static thread_local bool tl_stage_dep = false;    // host path: the next single-query scan launch depends programmatically on the staging pull
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU_OK(expr)                                                                        \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(e__ == cudaErrorMemoryAllocation ? MVDB_ERR_OOM : MVDB_ERR_CUDA,   \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                         \
    } while (0)
#define RC_OK(expr)             \
    do {                        \
        int rc__ = (expr);      \
        if (rc__ != MVDB_OK) return rc__; \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// driver entry points for virtual memory management, resolved at run time so
// that the library has no link-time dependency on libcuda (it must dlopen on
// a GPU-less build box for the symbol-export test).
// ---------------------------------------------------------------------------
struct DriverVmm {
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
        (void)cudaGetLastError();
    });
    return fn;
}

static DriverVmm* driver_vmm() {
    static DriverVmm api;
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("MVDB_NO_VMM")) return;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult st;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess &&
                   st == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        bool ok = true;
        ok &= get("cuMemAddressReserve", (void**)&api.AddressReserve);
        ok &= get("cuMemAddressFree", (void**)&api.AddressFree);
        ok &= get("cuMemCreate", (void**)&api.Create);
        ok &= get("cuMemRelease", (void**)&api.Release);
        ok &= get("cuMemMap", (void**)&api.Map);
        ok &= get("cuMemUnmap", (void**)&api.Unmap);
        ok &= get("cuMemSetAccess", (void**)&api.SetAccess);
        ok &= get("cuMemGetAllocationGranularity", (void**)&api.GetGranularity);
        (void)cudaGetLastError();
        api.ok = ok;
    });
    return &api;
}

// Growable device buffer.  VMM flavour: stable base address.  Fallback
// (MVDB_NO_VMM=1 or driver without VMM): cudaMalloc + copy, address may move
// (callers take the exclusive move lock around ensure()).
class GrowBuf {
  public:
    int init(int device, size_t reserve_bytes) {
        device_ = device;
        DriverVmm* v = driver_vmm();
        if (v->ok) {
            CUmemAllocationProp prop = {};
            prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
            prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            prop.location.id = device;
            size_t gran = 0;
            if (v->GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran) {
                gran_ = gran;
                reserved_ = align_up(std::max(reserve_bytes, gran), gran);
                if (v->AddressReserve(&base_, reserved_, 0, 0, 0) == CUDA_SUCCESS) {
                    vmm_ = true;
                    return MVDB_OK;
                }
            }
        }
        vmm_ = false;
        return MVDB_OK;
    }
    bool stable() const { return vmm_; }
    void* ptr() const { return vmm_ ? reinterpret_cast<void*>(base_) : plain_; }
    size_t mapped() const { return mapped_; }

    // Make at least `bytes` usable.  New bytes are zero-filled on `stream`.
    int ensure(size_t bytes, cudaStream_t stream) {
        if (bytes <= mapped_) return MVDB_OK;
        if (vmm_) {
            DriverVmm* v = driver_vmm();
            while (mapped_ < bytes) {
                // chunk doubles with the buffer, capped at 1 GiB, so handle
                // count stays small from kilobyte test indexes to 180 GB shards
                size_t want = std::min<size_t>(std::max(mapped_, gran_), size_t(1) << 30);
                want = std::max(want, std::min(bytes - mapped_, size_t(1) << 30));
                size_t chunk = align_up(want, gran_);
                if (mapped_ + chunk > reserved_ && bytes <= reserved_) chunk = reserved_ - mapped_;
                if (mapped_ + chunk > reserved_)
                    return fail(MVDB_ERR_OOM, "index capacity reservation (%zu bytes) exhausted", reserved_);
                CUmemAllocationProp prop = {};
                prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
                prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
                prop.location.id = device_;
                CUmemGenericAllocationHandle h;
                CUresult r = v->Create(&h, chunk, &prop, 0);
                if (r != CUDA_SUCCESS)
                    return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? MVDB_ERR_OOM : MVDB_ERR_CUDA,
                                "cuMemCreate(%zu) failed: %d", chunk, int(r));
                r = v->Map(base_ + mapped_, chunk, 0, h, 0);
                if (r != CUDA_SUCCESS) {
                    v->Release(h);
                    return fail(MVDB_ERR_CUDA, "cuMemMap failed: %d", int(r));
                }
                CUmemAccessDesc acc = {};
                acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
                acc.location.id = device_;
                acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
                r = v->SetAccess(base_ + mapped_, chunk, &acc, 1);
                if (r != CUDA_SUCCESS) {
                    v->Unmap(base_ + mapped_, chunk);
                    v->Release(h);
                    return fail(MVDB_ERR_CUDA, "cuMemSetAccess failed: %d", int(r));
                }
                CU_OK(cudaMemsetAsync(reinterpret_cast<void*>(base_ + mapped_), 0, chunk, stream));
                chunks_.push_back({h, chunk});
                mapped_ += chunk;
            }
            return MVDB_OK;
        }
        size_t cap = std::max(bytes, mapped_ * 2);
        cap = align_up(std::max<size_t>(cap, 1 << 16), 256);
        void* np = nullptr;
        CU_OK(cudaMalloc(&np, cap));
        CU_OK(cudaMemsetAsync(np, 0, cap, stream));
        if (plain_) {
            CU_OK(cudaMemcpyAsync(np, plain_, mapped_, cudaMemcpyDeviceToDevice, stream));
            CU_OK(cudaStreamSynchronize(stream));
            cudaFree(plain_);
        }
        plain_ = np;
        mapped_ = cap;
        return MVDB_OK;
    }
    void destroy() {
        if (vmm_) {
            DriverVmm* v = driver_vmm();
            size_t off = 0;
            for (auto& c : chunks_) {
                v->Unmap(base_ + off, c.bytes);
                v->Release(c.h);
                off += c.bytes;
            }
            chunks_.clear();
            if (base_) v->AddressFree(base_, reserved_);
            base_ = 0;
        } else if (plain_) {
            cudaFree(plain_);
            plain_ = nullptr;
        }
        mapped_ = 0;
    }

  private:
    struct Chunk {
        CUmemGenericAllocationHandle h;
        size_t bytes;
    };
    int device_ = 0;
    bool vmm_ = false;
    CUdeviceptr base_ = 0;
    size_t reserved_ = 0, mapped_ = 0, gran_ = 0;
    std::vector<Chunk> chunks_;
    void* plain_ = nullptr;
};

// simple grow-only device / pinned scratch
template <class T>
static int grow_dev(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return MVDB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    size_t nc = std::max(need, *cap * 2);
    CU_OK(cudaMalloc(reinterpret_cast<void**>(p), nc * sizeof(T)));
    *cap = nc;
    return MVDB_OK;
}
template <class T>
static int grow_pin(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return MVDB_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    size_t nc = std::max(need, *cap * 2);
    CU_OK(cudaMallocHost(reinterpret_cast<void**>(p), nc * sizeof(T)));
    *cap = nc;
    return MVDB_OK;
}

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
struct mvdb_workspace {
    mvdb_index* ix = nullptr;
    cudaStream_t stream = nullptr;  // owned (host-buffer searches run here)
    bool pooled = false;
    // device scratch
    uint64_t* partials = nullptr;
    size_t partials_cap = 0;
    unsigned int* ticket = nullptr;
    unsigned int launch_seq = 0;   // parity selects the tile counter of a scan launch
    bool prev_scan_big = false;    // the previous launch on this workspace was a one-CTA-per-SM scan (see "pdl")
    uint32_t* all_ord = nullptr;
    size_t all_ord_cap = 0;
    RadixState* radix = nullptr;
    FselState* fsel = nullptr;     // fast select for 128 < k <= 8192 on the host path (aux_kernels.cuh)
    uint64_t* keys = nullptr;
    size_t keys_cap = 0;
    // batched tensor-core path
    float* b_qn = nullptr;
    size_t b_qn_cap = 0;
    float* b_qnorm = nullptr;
    size_t b_qnorm_cap = 0;
    __nv_bfloat16* b_q16 = nullptr;
    size_t b_q16_cap = 0;
    uint64_t* b_thr = nullptr;
    size_t b_thr_cap = 0;
    unsigned int* b_cnt = nullptr;
    size_t b_cnt_cap = 0;
    unsigned int* b_ovf = nullptr;
    size_t b_ovf_cap = 0;
    uint64_t* b_cand = nullptr;
    size_t b_cand_cap = 0;
    const uint32_t** b_qmptr = nullptr;
    size_t b_qmptr_cap = 0;
    uint32_t* b_qmwords = nullptr;
    size_t b_qmwords_cap = 0;
    uint32_t* b_lmask = nullptr;      // sampled-row-space masks of the current level
    size_t b_lmask_cap = 0;
    uint32_t* b_lqmask = nullptr;
    size_t b_lqmask_cap = 0;
    const uint32_t** b_lqptr = nullptr;
    size_t b_lqptr_cap = 0;
    uint32_t* b_lqwords = nullptr;
    size_t b_lqwords_cap = 0;
    // survivor mode of the fp32 scan (32 < k <= 128)
    uint64_t* sv_surv = nullptr;
    unsigned int* sv_best = nullptr;
    SurvCtl* sv_ctl = nullptr;
    // int8 shadow mode (scan_i8.cuh)
    uint64_t* i8_surv = nullptr;
    unsigned int* i8_ovf_pin = nullptr;   // pinned + mapped: raised by the kernel when the survivor list overflowed
    unsigned int* i8_ovf_dev = nullptr;
    unsigned int* i8_best = nullptr;
    I8Ctl* i8_ctl = nullptr;
    // host-buffer path
    float* q_dev = nullptr;
    size_t q_cap = 0;
    uint32_t* mask_dev = nullptr;
    size_t mask_cap = 0;
    int64_t* I_dev = nullptr;   // [labels | distances] of a host-buffer search (one D2H transfer)
    size_t I_cap = 0;
    float* q_pin = nullptr;
    size_t q_pin_cap = 0;
    uint32_t* mask_pin = nullptr;
    size_t mask_pin_cap = 0;
    int64_t* I_pin = nullptr;
    size_t I_pin_cap = 0;
};

struct QMaskRef {          // one query's own admissible bitmask, resident on the device
    const uint32_t* dev;
    uint32_t words;        // 32-row words behind dev; rows past them are not admissible
};

struct mvdb_mask {         // device-resident filter, uploaded once, reusable by any number of searches
    mvdb_index* ix;
    uint32_t* dev;
    uint64_t rows;
    uint32_t words;
    // The mask kernels (predicate / fill / combine) run on the index's filter stream; `ready` is
    // recorded after the last one that wrote this mask and every search that reads the mask makes its
    // own (non-blocking) stream wait for it -- no host synchronisation between filter and search.
    cudaEvent_t ready = nullptr;
};

struct mvdb_column {       // one numeric metadata column, row-aligned with the index
    mvdb_index* ix;
    double* vals = nullptr;
    uint32_t* has = nullptr;
    uint64_t len = 0, cap = 0;   // rows
};

struct CoalesceReq {
    const float* q;
    int64_t k;
    const uint8_t* mask;
    uint64_t mask_rows;
    const mvdb_mask* handle = nullptr;
    int normalize;
    float* D;
    int64_t* I;
    int rc = MVDB_OK;
    bool done = false, promote = false;
    std::string err;
    std::condition_variable cv;
};

struct mvdb_index {
    int d = 0, device = 0;
    int64_t ld = 0;  // floats, multiple of 4
    int ld4 = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;
    GrowBuf mat, live;
    std::vector<uint32_t> live_host;  // mirror of the device bitmask
    std::atomic<uint64_t> ntotal{0};
    std::atomic<uint64_t> ndead{0};
    std::mutex mut_mu;            // serialises mutations
    std::shared_mutex move_mu;    // shared: searches / reads; exclusive: anything that moves rows
    cudaStream_t mut_stream = nullptr;
    float* stage_dev[2] = {nullptr, nullptr};
    size_t stage_cap[2] = {0, 0};
    int64_t* rows_dev = nullptr;
    size_t rows_cap = 0;
    // bf16 shadow of the matrix for the tensor-core batched path (built lazily, kept in step by
    // converting rows [shadow_rows, ntotal) at the start of a batched search)
    GrowBuf mat16;
    int64_t ld16 = 0;              // bf16 elements per shadow row (multiple of 8)
    uint64_t shadow_rows = 0;
    std::mutex shadow_mu;
    // int8 shadow of the matrix for the opt-in low-precision single-query mode (option "scan_shadow"; built lazily)
    GrowBuf mat8;
    int ld8 = 0;                    // int8 elements per record (d rounded up to 16)
    uint32_t rec8 = 0;              // record stride in bytes (ld8 + 16: scale, residual norm, pad)
    uint64_t shadow8_rows = 0;
    std::mutex shadow8_mu;
    int scan_shadow = 0;
    int survivor_tail = 1;          // 32 < k <= 128 single-query scans: shared threshold + global survivor list (0 = per-warp selects + merge tree)
    int* max_norm2_bits = nullptr;  // device: bit pattern of the largest squared row norm stored
    std::atomic<float> max_norm2_host{0.f};  // host copy, refreshed at the end of every add
    // options
    int batch_mode = 1;            // 0 off, 1 exact (bf16 candidates + fp32 re-scoring), 2 bf16, 3 tf32
    int batch_min_nq = 2;
    int batch_cost_model = 1;      // 0: every batch of >= batch_min_nq queries takes the tensor path (tests, probes)
    int gemm_l2_hint = 0;
    unsigned long long* trace_dev = nullptr;   // debug timeline of the scan kernel (option "trace")
    int gemm_short_a = 1;          // batches of <= 64 queries stage a short A tile and run a deeper ring (0 = always 128 rows)
    int gemm_debug = 0;            // GemmParams::debug experiments (results are garbage when non-zero)
    unsigned long long* gemm_prof_dev = nullptr;   // debug wait-cycle counters of the GEMM kernels (option "gemm_prof"), [256][8]
    int pdl = 0;                   // search_device: programmatic dependent launch of back-to-back scans (opt-in)
    int large_k_fast = 1;          // host-buffer searches with 128 < k <= 8192: histogram select (5 launches) instead of the radix select (22)
    int host_path = 3;             // host-buffer single queries: 1 results written straight to pinned host memory, 2 inputs
                                   // pulled by a grid the scan depends on programmatically
    int dyn_tiles = 15;            // % of the tiles the TMA scan claims from a global counter (rest: static round-robin)
    int l2_pin_mb = 0;             // head of the matrix kept L2-resident across scans (evict_last), MB
    int gemm_variant = 2;          // 0: one CTA per 128x256 tile; 1: CTA pairs (cta_group::2), 256x256 tiles;
                                   // 2 / 3: cluster of 2 / 4 CTAs sharing the X tile through TMA multicast
    int scan_variant = MVDB_SCAN_AUTO;
    int fused_k_max = 128;
    int grid_ctas = 0;
    int consumer_warps = 0;
    unsigned long long* count_dev = nullptr;   // scratch of mvdb_mask_count
    std::mutex count_mu;
    cudaStream_t filter_stream = nullptr;      // every device-side filter kernel runs here, in call order
    // query coalescer: concurrent single-query host searches share one pass over the matrix
    int coalesce = 1;
    int coalesce_max = 64;
    std::mutex co_mu;
    std::deque<struct CoalesceReq*> co_queue;
    int co_leaders = 0;
    int co_max_leaders = 0;        // option "coalesce_leaders": 0 = auto (1 for large matrices, else 2)
    size_t co_last_batch = 0;      // sizes of the last two batches: how much company a new leader may expect (their SUM --
    size_t co_prev_batch = 0;      // a round of callers that split into two batches must be able to merge again)
    int co_wait_pct = 40;          // option "coalesce_wait_pct": a leader waits at most this % of a pass for that company (0 = never)
    // objects that point back at this index (mask handles, columns, caller-owned workspaces): destroying the
    // index releases their device memory and orphans them, so that a later *_destroy of theirs (e.g. from a
    // garbage collector that runs after the index is gone) is harmless
    std::mutex child_mu;
    std::unordered_set<mvdb_mask*> child_masks;
    std::unordered_set<mvdb_column*> child_columns;
    std::unordered_set<mvdb_workspace*> child_workspaces;
    struct mvdb_group* group = nullptr;   // the shard group this index belongs to (it must be destroyed first)
    // workspace pool for host-buffer searches
    std::mutex pool_mu;
    std::condition_variable pool_cv;
    std::vector<mvdb_workspace*> pool_free;
    int pool_created = 0;
    static constexpr int kPoolMax = 16;
};

struct mvdb_exchange {
    int device = 0, rank = 0, world = 1, k_max = 0, nq_max = 0;
    uint64_t* local = nullptr;      // [recv | flags], cudaMalloc (IPC-exportable)
    size_t recv_words = 0;
    void* peer_base[kMaxWorld] = {};
    XchgDev host = {};
    XchgDev* dev = nullptr;
    unsigned int* status = nullptr;        // pinned, mapped host word written by the kernel on a peer timeout
    unsigned int* status_dev = nullptr;    // its device alias
    uint64_t timeout_ns = 2000000000ull;   // option "timeout_ms" (default 2 s)
    uint64_t seq = 0;
    bool connected = false;
    bool local_peers = false;              // connected through same-process peer access (nothing to unmap)
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define ENTER(ix)                                                                    \
    if (!(ix)) return fail(MVDB_ERR_ARG, "null index");                              \
    DeviceGuard guard__((ix)->device);                                               \
    if (!guard__.ok) return fail(MVDB_ERR_CUDA, "cudaSetDevice(%d) failed", (ix)->device); \
    (void)cudaGetLastError() /* start from a clean per-thread error state: cudaGetLastError() below must report OUR launches */

// ---------------------------------------------------------------------------
// scan dispatch
// ---------------------------------------------------------------------------
typedef void (*ScanKernel)(const ScanParams);

template <bool kTma>
static ScanKernel q1_kernel(int d4) {
    switch (d4) {
        case 1: return scan_q1_kernel<1, kTma>;
        case 2: return scan_q1_kernel<2, kTma>;
        case 3: return scan_q1_kernel<3, kTma>;
        case 4: return scan_q1_kernel<4, kTma>;
        case 5: return scan_q1_kernel<5, kTma>;
        case 6: return scan_q1_kernel<6, kTma>;
        case 7: return scan_q1_kernel<7, kTma>;
        case 8: return scan_q1_kernel<8, kTma>;
        default: return nullptr;
    }
}
static ScanKernel q1_survivor_kernel(int d4) {
    switch (d4) {
        case 1: return scan_q1_kernel<1, true, true>;
        case 2: return scan_q1_kernel<2, true, true>;
        case 3: return scan_q1_kernel<3, true, true>;
        case 4: return scan_q1_kernel<4, true, true>;
        case 5: return scan_q1_kernel<5, true, true>;
        case 6: return scan_q1_kernel<6, true, true>;
        case 7: return scan_q1_kernel<7, true, true>;
        case 8: return scan_q1_kernel<8, true, true>;
        default: return nullptr;
    }
}
template <bool kTma>
static ScanKernel multi_kernel(int nq) {
    switch (nq) {
        case 1: return scan_multi_kernel<1, kTma>;
        case 2: return scan_multi_kernel<2, kTma>;
        case 4: return scan_multi_kernel<4, kTma>;
        case 8: return scan_multi_kernel<8, kTma>;
        default: return nullptr;
    }
}

struct ScanPlan {
    ScanKernel fn = nullptr;
    bool tma = false;
    bool q1 = false;
    int grid = 0, threads = 0;
    size_t smem = 0;
};

// Choose kernel, grid and shared-memory layout for one launch of `nq`
// (1, 2, 4 or 8) queries.  Fills the layout fields of `p`.
static int plan_scan(mvdb_index* ix, ScanParams& p, int nq, ScanPlan* plan) {
    const int d4 = (ix->ld4 + 31) / 32;
    const bool use_q1 = (nq == 1 && d4 <= 8);
    const uint32_t tiles = (p.n + kRowsPerTile - 1) / kRowsPerTile;
    p.cap = select_cap(p.k);
    p.nq = nq;
    p.stage_bytes = uint32_t(align_up(size_t(kRowsPerTile) * ix->ld * 4, 128));

    auto layout = [&](int ncw, bool tma, int* stages) -> size_t {
        size_t off = 1024;
        p.sel_off = uint32_t(off);
        off += size_t(ncw) * nq * p.cap * 8;
        off = align_up(off, 16);
        p.q_off = uint32_t(off);
        if (!use_q1) off += size_t(nq) * ix->ld * 4;
        off = align_up(off, 128);
        p.stage_off = uint32_t(off);
        if (!tma) {
            // LDG variant: a tail region only for the last-CTA merge scratch
            *stages = 0;
            size_t want = std::min<size_t>(size_t(ix->sm_count) * 4 * p.k * 8, 32 * 1024);
            p.merge_off = uint32_t(off);
            p.merge_bytes = uint32_t(want);
            return off + want;
        }
        if (off + 2 * size_t(p.stage_bytes) > ix->smem_optin) {
            *stages = 0;
            return 0;
        }
        int s = int((ix->smem_optin - off) / p.stage_bytes);
        s = std::min(s, 16);
        *stages = s;
        return off + size_t(s) * p.stage_bytes;
    };

    // TMA ring: stage s of the ring must always be consumed by the same warp
    // (consumer of iteration `it` is warp it % ncw, its stage is it % S), else a
    // warp that runs ahead could pass a full-barrier wait on the parity of an
    // OLDER phase.  So S is rounded down to a multiple of ncw, and ncw is
    // chosen to keep the ring deep for wide rows (e.g. d=1024: 7 stages x 7 warps).
    int variant = ix->scan_variant;
    int ncw_tma = 0, stages = 0;
    size_t smem = 0;
    bool tma = (variant != MVDB_SCAN_LDG);
    if (tma) {
        const int pref = nq >= 4 ? 8 : 4;
        int cand[9] = {pref, 8, 7, 6, 5, 4, 3, 2, 0};
        if (ix->consumer_warps > 0) {
            cand[0] = std::min(ix->consumer_warps, 8);
            cand[1] = 0;
        }
        int best_score = 0;
        for (int i = 0; cand[i]; i++) {
            int s_raw = 0;
            layout(cand[i], true, &s_raw);
            int s_ok = s_raw / cand[i] * cand[i];
            int score = std::min(s_ok, 8);
            if (score > best_score) {
                best_score = score;
                ncw_tma = cand[i];
                stages = s_ok;
            }
        }
        // the ring needs >= 2 stages to overlap at all; AUTO wants >= 3
        if (stages < 2 || (stages < 3 && variant != MVDB_SCAN_TMA)) tma = false;
        if (tma) {
            int s_raw = 0;
            layout(ncw_tma, true, &s_raw);  // re-derive the offsets for the chosen ncw
            smem = size_t(p.stage_off) + size_t(stages) * p.stage_bytes;
            p.merge_off = p.stage_off;     // the ring is idle once every tile is consumed
            p.merge_bytes = uint32_t(size_t(stages) * p.stage_bytes);
        }
    }
    plan->q1 = use_q1;
    if (tma) {
        plan->tma = true;
        plan->threads = 32 * (1 + ncw_tma);
        plan->fn = use_q1 ? q1_kernel<true>(d4) : multi_kernel<true>(nq);
        plan->smem = smem;
        p.stages = stages;
        int g = ix->grid_ctas > 0 ? ix->grid_ctas : ix->sm_count;
        plan->grid = int(std::min<uint32_t>(uint32_t(g), tiles));
    } else {
        const int ncw = 8;
        smem = layout(ncw, false, &stages);
        if (smem > ix->smem_optin) return fail(MVDB_ERR_ARG, "k=%d / d=%d need too much shared memory", p.k, ix->d);
        plan->tma = false;
        plan->threads = 32 * ncw;
        plan->fn = use_q1 ? q1_kernel<false>(d4) : multi_kernel<false>(nq);
        plan->smem = smem;
        p.stages = 0;
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan->fn, plan->threads, smem);
        per_sm = std::max(1, std::min(per_sm, 4));
        int g = ix->grid_ctas > 0 ? ix->grid_ctas : ix->sm_count * per_sm;
        uint32_t per_cta_tiles = uint32_t(ncw);  // at least one tile per warp
        plan->grid = int(std::max<uint32_t>(1, std::min<uint32_t>(uint32_t(g), (tiles + per_cta_tiles - 1) / per_cta_tiles)));
    }
    if (!plan->fn) return fail(MVDB_ERR_STATE, "no scan kernel for nq=%d d4=%d", nq, d4);
    static std::mutex attr_mu;
    {
        // opt in to large dynamic shared memory once per kernel
        std::lock_guard<std::mutex> g(attr_mu);
        static std::vector<std::pair<void*, int>> done;
        bool seen = false;
        for (auto& e : done) seen |= (e.first == (void*)plan->fn && e.second == ix->device);
        if (!seen) {
            CU_OK(cudaFuncSetAttribute(plan->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ix->smem_optin)));
            done.push_back({(void*)plan->fn, ix->device});
        }
    }
    return MVDB_OK;
}

__global__ void fill_empty_results_kernel(float* D, int64_t* I, int64_t total) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < total) {
        D[i] = -FLT_MAX;
        I[i] = -1;
    }
}

static int ws_scratch(mvdb_workspace* ws) {
    if (!ws->ticket) {
        // ticket at +0, the dynamic scheduler's two tile counters at +128 and +256 (own L2 lines)
        CU_OK(cudaMalloc(&ws->ticket, 512));
        CU_OK(cudaMemset(ws->ticket, 0, 512));
    }
    return MVDB_OK;
}


// ---------------------------------------------------------------------------
// batched tensor-core path (gemm_tc.cuh)
// ---------------------------------------------------------------------------
// rows x d bf16 matrix whose consecutive (sampled) rows are ld_elems elements apart
// tf32 = false: bf16 elements, 64 per 128-byte box row; tf32 = true: fp32 elements, 32 per box row
static int encode_gemm_map(CUtensorMap* tm, const void* base, uint64_t rows, int d, int64_t ld_elems, uint32_t box_rows,
                           bool tf32 = false) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(MVDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const unsigned esize = tf32 ? 4u : 2u;
    cuuint64_t gdim[2] = {cuuint64_t(d), cuuint64_t(std::max<uint64_t>(rows, 1))};
    cuuint64_t gstride[1] = {cuuint64_t(ld_elems) * esize};
    cuuint32_t box[2] = {cuuint32_t(128u / esize), box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = enc(tm, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                     gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MVDB_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", int(r));
    return MVDB_OK;
}

// bring the bf16 shadow up to `n` rows (round-to-nearest conversion of the stored fp32 rows)
static int ensure_shadow(mvdb_index* ix, uint64_t n) {
    std::lock_guard<std::mutex> g(ix->shadow_mu);
    if (ix->shadow_rows >= n) return MVDB_OK;
    cudaStream_t st = ix->mut_stream;
    RC_OK(ix->mat16.ensure(size_t(n) * ix->ld16 * 2, st));
    const uint64_t r0 = ix->shadow_rows, m = n - r0;
    unsigned grid = unsigned(std::min<uint64_t>((m * 32 + 255) / 256, uint64_t(ix->sm_count) * 16));
    to_bf16_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(ix->mat.ptr()) + r0 * ix->ld,
                                              static_cast<__nv_bfloat16*>(ix->mat16.ptr()) + r0 * ix->ld16, m, ix->d, ix->ld,
                                              ix->ld16);
    LAUNCHED();
    CU_OK(cudaGetLastError());
    CU_OK(cudaStreamSynchronize(st));
    ix->shadow_rows = n;
    return MVDB_OK;
}

static constexpr uint32_t kCandCap = 8192;      // candidate slots per query
static constexpr uint32_t kFirstLevel = 2048;   // sampled rows of the coarsest level (all admissible ones become candidates)

static int run_search(mvdb_index* ix, mvdb_workspace* ws, const float* q_dev, int64_t nq, int64_t k,
                      const uint32_t* mask_dev, uint64_t mask_rows, int normalize_q, int64_t label_offset,
                      float* D_dev, int64_t* I_dev, cudaStream_t stream, mvdb_exchange* xch,
                      const QMaskRef* qmasks = nullptr);

// Q[nq,d] against rows [0,n): bf16 GEMM on tcgen05 with threshold-filter epilogue over
// multi-resolution strided samples (thresholds tighten level by level), optional exact re-scoring.
// dense_out != nullptr: debug mode, write every bf16-GEMM score to dense_out[nq][n].
static int run_batched(mvdb_index* ix, mvdb_workspace* ws, const float* q_dev, int64_t nq, int64_t k,
                       const uint32_t* mask_dev, uint32_t n, int normalize_q, int64_t label_offset, float* D_dev,
                       int64_t* I_dev, cudaStream_t stream, int mode, float* dense_out,
                       const QMaskRef* qmasks = nullptr) {
    // mode 1 exact (bf16 candidates + fp32 re-score), 2 bf16 scores, 3 tf32 scores (fp32 operands
    // straight from the master matrix: no shadow copy, half the tensor rate, twice the bytes)
    const bool tf32 = mode == 3;
    ws->prev_scan_big = false;
    if (!tf32) RC_OK(ensure_shadow(ix, n));
    RC_OK(grow_dev(&ws->b_qn, &ws->b_qn_cap, size_t(nq) * ix->ld));
    RC_OK(grow_dev(&ws->b_qnorm, &ws->b_qnorm_cap, size_t(nq)));
    RC_OK(grow_dev(&ws->b_q16, &ws->b_q16_cap, size_t(nq) * ix->ld16));
    RC_OK(grow_dev(&ws->b_thr, &ws->b_thr_cap, size_t(nq)));
    RC_OK(grow_dev(&ws->b_cnt, &ws->b_cnt_cap, size_t(nq)));
    RC_OK(grow_dev(&ws->b_ovf, &ws->b_ovf_cap, size_t(nq)));
    if (!dense_out) RC_OK(grow_dev(&ws->b_cand, &ws->b_cand_cap, size_t(nq) * kCandCap));

    prep_queries_kernel<<<unsigned((nq * 32 + 255) / 256), 256, 0, stream>>>(q_dev, ws->b_qn, ws->b_qnorm, nq, ix->d, ix->ld,
                                                                            normalize_q);
    LAUNCHED();
    if (!tf32) {
        to_bf16_rows_kernel<<<unsigned(std::min<int64_t>((nq * 32 + 255) / 256, 4096)), 256, 0, stream>>>(
            ws->b_qn, ws->b_q16, uint64_t(nq), ix->d, ix->ld, ix->ld16);
        LAUNCHED();
    }
    init_batch_state_kernel<<<unsigned((nq + 255) / 256), 256, 0, stream>>>(ws->b_thr, ws->b_cnt, ws->b_ovf, nq);
    LAUNCHED();

    CUtensorMap tmQ, tmX, tmQ2, tmX2, tmX4;
    const void* const x_base = tf32 ? ix->mat.ptr() : ix->mat16.ptr();
    const int64_t x_ld = tf32 ? int64_t(ix->ld) : int64_t(ix->ld16);
    const int variant = tf32 ? (ix->gemm_variant >= 2 ? 2 : 0) : ix->gemm_variant;   // tf32: single CTA or cluster-2 multicast
    // one query block (nq <= 128) always runs the single-CTA kernel: stage only as many query rows as
    // there are (32 / 64 / 128) and spend the shared memory on a deeper ring (6 / 5 / 4 stages)
    const uint32_t a_rows = (nq <= 32 && ix->gemm_short_a) ? 32u : (nq <= 64 && ix->gemm_short_a) ? 64u : uint32_t(kGemmBM);
    const uint32_t v0_stages = a_rows == 32 ? 6u : a_rows == 64 ? 5u : uint32_t(kGemmStages);
    const size_t v0_smem = size_t(v0_stages) * (a_rows * 128u + kGemmBBytes) + 256 + 1024;
    if (tf32) RC_OK(encode_gemm_map(&tmQ, ws->b_qn, uint64_t(nq), ix->d, ix->ld, a_rows, true));
    else RC_OK(encode_gemm_map(&tmQ, ws->b_q16, uint64_t(nq), ix->d, ix->ld16, a_rows));
    tmQ2 = tmQ;   // the pair kernel loads 128-query boxes too
    tmX2 = tmQ;
    tmX4 = tmQ;
    CU_OK(cudaFuncSetAttribute(gemm_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ix->smem_optin)));
    CU_OK(cudaFuncSetAttribute(gemm_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ix->smem_optin)));
    CU_OK(cudaFuncSetAttribute(gemm_topk_kernel_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGemm2SmemBytes)));
    CU_OK(cudaFuncSetAttribute(gemm_topk_kernel_mc<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGemmSmemBytes)));
    CU_OK(cudaFuncSetAttribute(gemm_topk_kernel_mc<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGemmSmemBytes)));
    CU_OK(cudaFuncSetAttribute(gemm_topk_kernel_mc<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGemmSmemBytes)));
    CU_OK(cudaFuncSetAttribute(cand_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCandCap * 8)));

    // rigorous bound on |bf16 score - fp32 score| per unit |q|: inputs rounded to nearest
    // bf16 (2^-9 relative each) + fp32 accumulation of d terms, times the largest row norm
    float slack_unit = 0.f;
    if (mode == 1 && !dense_out) {
        const float max_norm2 = ix->max_norm2_host.load(std::memory_order_acquire);
        const double eps_unit = (std::ldexp(1.0, -8) + std::ldexp(1.0, -18) + double(ix->d) * std::ldexp(1.0, -22)) * 1.01;
        slack_unit = float(2.0 * eps_unit * std::sqrt(std::max(double(max_norm2), 1e-30)) * 1.0001);
    }

    const uint32_t* live_dev = ix->ndead.load(std::memory_order_acquire) ? static_cast<const uint32_t*>(ix->live.ptr()) : nullptr;
    GemmParams gp = {};
    if (qmasks) {
        // per-query filters: thread <-> query in the epilogue, so each thread reads its own words
        std::vector<const uint32_t*> ptrs(size_t(nq), nullptr);
        std::vector<uint32_t> wrds(size_t(nq), 0u);
        for (int64_t i = 0; i < nq; i++) {
            ptrs[size_t(i)] = qmasks[i].dev;
            wrds[size_t(i)] = qmasks[i].words;
        }
        RC_OK(grow_dev(&ws->b_qmptr, &ws->b_qmptr_cap, size_t(nq)));
        RC_OK(grow_dev(&ws->b_qmwords, &ws->b_qmwords_cap, size_t(nq)));
        CU_OK(cudaMemcpyAsync(ws->b_qmptr, ptrs.data(), size_t(nq) * 8, cudaMemcpyHostToDevice, stream));
        CU_OK(cudaMemcpyAsync(ws->b_qmwords, wrds.data(), size_t(nq) * 4, cudaMemcpyHostToDevice, stream));
        CU_OK(cudaStreamSynchronize(stream));   // the host vectors go out of scope
    }
    gp.nq = nq;
    gp.d = ix->d;
    gp.l2_hint = ix->gemm_l2_hint;
    gp.thr = ws->b_thr;
    gp.cand = ws->b_cand;
    gp.cand_cnt = ws->b_cnt;
    gp.cand_cap = kCandCap;
    gp.dense = dense_out;
    gp.prof = ix->gemm_prof_dev;
    gp.debug = ix->gemm_debug;
    gp.a_rows = a_rows;
    gp.stages = v0_stages;
    gp.dense_ld = n;
    const uint32_t n_qb = uint32_t((nq + kGemmBM - 1) / kGemmBM);

    // Multi-resolution strided sampling.  Level strides are powers of 16 down to 1: the
    // coarsest level holds <= kFirstLevel rows spread evenly over the whole index (every
    // admissible one becomes a candidate and fixes a first threshold per query), each finer
    // level contains the previous one (so its threshold stays a valid lower bound), restarts
    // its candidate list and admits ~16 k rows per query; the last level is the full matrix.
    // Unlike contiguous chunks this does not depend on the ORDER of the rows: a deleted or
    // filtered-out prefix (time-ordered data with a date filter, churn that deletes the oldest
    // rows) cannot flood the candidate lists.  Extra flops: 1/16 + 1/256 + ... = 6.7 %.
    std::vector<uint32_t> strides;
    // Small batches (one query block) on mid-size matrices are dominated by launch count, not by
    // candidate volume: two levels are enough when the final level admits <= ~2048 rows per query
    // (k * stride), and the final level then runs in one piece.
    bool single_piece = false;
    {
        uint32_t s2 = 1;
        while ((uint64_t(n) + s2 - 1) / s2 > 8192 - 256) s2 *= 2;
        if (!dense_out && nq <= kGemmBM && s2 > 1 && uint64_t(k) * s2 <= 2048) {
            strides = {s2, 1};
            single_piece = true;
        } else {
            uint32_t s0 = 1;
            while (!dense_out && (uint64_t(n) + s0 - 1) / s0 > kFirstLevel) s0 *= 16;
            for (uint32_t st_ = s0;; st_ /= 16) {
                strides.push_back(st_);
                if (st_ == 1) break;
            }
        }
    }
    for (size_t lv = 0; lv < strides.size(); lv++) {
        const uint32_t S = strides[lv];
        const uint32_t m = uint32_t((uint64_t(n) + S - 1) / S);   // sampled rows of this level
        const uint32_t words = (m + 31) / 32;
        RC_OK(encode_gemm_map(&tmX, x_base, m, ix->d, x_ld * int64_t(S), kGemmBN, tf32));
        if (variant >= 1) RC_OK(encode_gemm_map(&tmX2, x_base, m, ix->d, x_ld * int64_t(S), kGemmBN / 2, tf32));
        if (variant == 3) RC_OK(encode_gemm_map(&tmX4, x_base, m, ix->d, x_ld * int64_t(S), kGemmBN / 4, tf32));
        gp.row_stride = S;
        gp.row0 = 0;
        gp.row1 = uint32_t(align_up(m, kGemmBN));
        gp.n_valid = m;
        gp.live = live_dev;
        gp.mask = mask_dev;
        gp.qmask = qmasks ? ws->b_qmptr : nullptr;
        gp.qmask_words = qmasks ? ws->b_qmwords : nullptr;
        if (S > 1) {
            // bring the bitmasks into this level's sampled-row space
            if (live_dev || mask_dev) {
                RC_OK(grow_dev(&ws->b_lmask, &ws->b_lmask_cap, size_t(words)));
                sample_mask_kernel<<<(words + 255) / 256, 256, 0, stream>>>(live_dev, mask_dev, n, S, ws->b_lmask, m);
                LAUNCHED();
                gp.live = nullptr;
                gp.mask = ws->b_lmask;
            }
            if (qmasks) {
                RC_OK(grow_dev(&ws->b_lqmask, &ws->b_lqmask_cap, size_t(words) * nq));
                RC_OK(grow_dev(&ws->b_lqptr, &ws->b_lqptr_cap, size_t(nq)));
                RC_OK(grow_dev(&ws->b_lqwords, &ws->b_lqwords_cap, size_t(nq)));
                sample_qmask_kernel<<<dim3((words + 255) / 256, unsigned(nq)), 256, 0, stream>>>(ws->b_qmptr, ws->b_qmwords, S,
                                                                                              ws->b_lqmask, m, words);
                LAUNCHED();
                qmask_table_kernel<<<unsigned((nq + 255) / 256), 256, 0, stream>>>(ws->b_qmptr, ws->b_lqmask, words, ws->b_lqptr,
                                                                                 ws->b_lqwords, nq);
                LAUNCHED();
                gp.qmask = ws->b_lqptr;
                gp.qmask_words = ws->b_lqwords;
            }
        }
        if (lv > 0) {
            reset_counts_kernel<<<unsigned((nq + 255) / 256), 256, 0, stream>>>(ws->b_cnt, nq);
            LAUNCHED();
        }
        // The full-matrix level runs as 4 contiguous pieces (1/8, 1/8, 1/4, 1/2) with a threshold
        // refresh after each: the sampled levels already bound what can pass (<= ~16 k per
        // query whatever the row order), the refreshes cut that to ~5 k.
        std::vector<uint32_t> cuts;
        const uint32_t rows_al = gp.row1;
        if (S == 1 && strides.size() > 1 && rows_al >= 64 * kGemmBN && !single_piece) {
            const uint32_t e = uint32_t(align_up(rows_al / 8, kGemmBN));
            cuts = {e, 2 * e, 4 * e, rows_al};
        } else {
            cuts = {rows_al};
        }
        uint32_t lo = 0;
        for (uint32_t hi : cuts) {
            gp.row0 = lo;
            gp.row1 = hi;
            if (variant >= 2 && nq > kGemmBM) {
                // clusters of CS CTAs share one X tile through TMA multicast (CS query blocks at once)
                const int cs = (variant == 3 && nq > 2 * kGemmBM) ? 4 : 2;
                const uint64_t tiles2 = uint64_t((hi - lo) / kGemmBN) * ((nq + cs * kGemmBM - 1) / (cs * kGemmBM));
                const unsigned clusters = unsigned(std::min<uint64_t>(uint64_t(ix->sm_count / cs), tiles2));
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(clusters * cs);
                cfg.blockDim = dim3(384);
                cfg.dynamicSmemBytes = kGemmSmemBytes;
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = unsigned(cs);
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                if (cs == 4) CU_OK(cudaLaunchKernelEx(&cfg, gemm_topk_kernel_mc<4, false>, tmQ2, tmX4, gp));
                else if (tf32) CU_OK(cudaLaunchKernelEx(&cfg, gemm_topk_kernel_mc<2, true>, tmQ2, tmX2, gp));
                else CU_OK(cudaLaunchKernelEx(&cfg, gemm_topk_kernel_mc<2, false>, tmQ2, tmX2, gp));
            } else if (variant == 1 && nq > kGemmBM) {
                // CTA pairs: 256-query x 256-row tiles, B operand split across the pair
                const uint64_t tiles2 = uint64_t((hi - lo) / kGemmBN) * ((nq + 2 * kGemmBM - 1) / (2 * kGemmBM));
                const unsigned pairs = unsigned(std::min<uint64_t>(uint64_t(ix->sm_count / 2), tiles2));
                gemm_topk_kernel_2cta<<<2 * pairs, 384, kGemm2SmemBytes, stream>>>(tmQ2, tmX2, gp);
            } else {
                const uint64_t tiles = uint64_t((hi - lo) / kGemmBN) * n_qb;
                const unsigned grid = unsigned(std::min<uint64_t>(uint64_t(ix->sm_count), tiles));
                if (tf32) gemm_topk_kernel<true><<<grid, 384, v0_smem, stream>>>(tmQ, tmX, gp);
                else gemm_topk_kernel<false><<<grid, 384, v0_smem, stream>>>(tmQ, tmX, gp);
            }
            LAUNCHED();
            if (!dense_out) {
                cand_update_kernel<<<unsigned(nq), 256, kCandCap * 8, stream>>>(ws->b_cand, ws->b_cnt, ws->b_thr, ws->b_ovf, kCandCap,
                                                                               int(k), mode == 1 ? ws->b_qnorm : nullptr, slack_unit);
                LAUNCHED();
            }
            lo = hi;
        }
    }
    CU_OK(cudaGetLastError());
    if (dense_out) return MVDB_OK;
    if (mode == 1) {
        rescore_kernel<<<unsigned(nq), 256, size_t(ix->ld) * 4, stream>>>(static_cast<const float*>(ix->mat.ptr()), ix->ld4, ws->b_qn,
                                                                         ws->b_cand, ws->b_cnt, kCandCap);
        LAUNCHED();
        cand_update_kernel<<<unsigned(nq), 256, kCandCap * 8, stream>>>(ws->b_cand, ws->b_cnt, ws->b_thr, ws->b_ovf, kCandCap, int(k),
                                                                       nullptr, 0.f);
        LAUNCHED();
    }
    batch_results_kernel<<<unsigned((nq * k + 255) / 256), 256, 0, stream>>>(ws->b_cand, ws->b_cnt, kCandCap, int(k), nq, label_offset,
                                                                            D_dev, I_dev);
    LAUNCHED();
    CU_OK(cudaGetLastError());
    // queries whose candidate list overflowed (adversarial row order) are redone by the exact scan
    std::vector<unsigned int> ovf(size_t(nq), 0u);
    CU_OK(cudaMemcpyAsync(ovf.data(), ws->b_ovf, size_t(nq) * 4, cudaMemcpyDeviceToHost, stream));
    CU_OK(cudaStreamSynchronize(stream));
    int rc = MVDB_OK;
    for (int64_t q = 0; q < nq && rc == MVDB_OK; q++) {
        if (!ovf[size_t(q)]) continue;
        tl_force_scan = true;
        rc = run_search(ix, ws, q_dev + q * ix->d, 1, k, mask_dev, n, normalize_q, label_offset, D_dev + q * k, I_dev + q * k,
                        stream, nullptr, qmasks ? qmasks + q : nullptr);
        tl_force_scan = false;
    }
    return rc;
}

// bring the int8 shadow up to `n` rows
static int ensure_shadow8(mvdb_index* ix, uint64_t n) {
    std::lock_guard<std::mutex> g(ix->shadow8_mu);
    if (ix->shadow8_rows >= n) return MVDB_OK;
    cudaStream_t st = ix->mut_stream;
    RC_OK(ix->mat8.ensure(size_t(n) * ix->rec8, st));
    const uint64_t r0 = ix->shadow8_rows, m = n - r0;
    unsigned grid = unsigned(std::min<uint64_t>((m * 32 + 255) / 256, uint64_t(ix->sm_count) * 16));
    to_i8_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(ix->mat.ptr()) + r0 * ix->ld,
                                            static_cast<uint8_t*>(ix->mat8.ptr()) + r0 * ix->rec8, m, ix->ld4, ix->ld8, ix->rec8);
    LAUNCHED();
    CU_OK(cudaGetLastError());
    CU_OK(cudaStreamSynchronize(st));
    ix->shadow8_rows = n;
    return MVDB_OK;
}

// consumer warps of the int8 scan: 8 when the ring can be >= 8 tiles deep, else 4 (rows wider than ~800 bytes)
static int i8_consumer_warps(const mvdb_index* ix) {
    const size_t q_bytes = align_up(size_t(ix->ld) * 4, 128);
    const size_t stage = align_up(size_t(kI8TileRows) * ix->rec8, 128);
    const size_t raw = (ix->smem_optin - 1024 - q_bytes) / stage;
    return raw >= 8 ? 8 : raw >= 4 ? 4 : 0;
}

// Can this search take the int8 shadow mode?  One query, fused-k range, rows narrow enough for a >= 4-deep
// ring of 32-record tiles, and enough rows for the shadow pass to beat the fp32 scan's fixed costs.
static bool i8_eligible(const mvdb_index* ix, int64_t nq, int64_t k, uint32_t n) {
    if (!ix->scan_shadow || nq != 1 || k > 128 || k > ix->fused_k_max || n < 16384 || ix->d > 1024) return false;
    return i8_consumer_warps(ix) != 0;
}

static constexpr uint32_t kSurvCap = 8192;   // = kSelectMax (scan.cuh)

// Launch with the programmatic-stream-serialisation attribute: the grid may start while its predecessor on the stream
// still runs; the kernel itself waits (griddepcontrol.wait) before it touches what the predecessor produces.
template <class P>
static cudaError_t launch_dependent(void (*fn)(P), int grid, int threads, size_t smem, cudaStream_t stream, const P& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(threads));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fn, p);
}

// pinned + mapped word a kernel raises when its candidate / survivor list overflowed (host-buffer callers check it
// after their synchronise and re-run the query on the classic scan)
static int ws_overflow_flag(mvdb_workspace* ws) {
    if (ws->i8_ovf_pin) return MVDB_OK;
    CU_OK(cudaHostAlloc(reinterpret_cast<void**>(&ws->i8_ovf_pin), sizeof(unsigned int), cudaHostAllocMapped));
    *ws->i8_ovf_pin = 0u;
    CU_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ws->i8_ovf_dev), ws->i8_ovf_pin, 0));
    return MVDB_OK;
}

// scratch of the survivor mode (allocations synchronise the device: a shard group does this before its first launch)
static int sv_prepare(mvdb_index* ix, mvdb_workspace* ws, int d4, cudaStream_t stream) {
    RC_OK(ws_overflow_flag(ws));
    if (!ws->sv_ctl) {
        CU_OK(cudaMalloc(&ws->sv_ctl, sizeof(SurvCtl)));
        CU_OK(cudaMalloc(&ws->sv_surv, size_t(kSurvCap) * 8));
        CU_OK(cudaMalloc(&ws->sv_best, size_t(kI8BestM) * 32 * 4));
        CU_OK(cudaMemsetAsync(ws->sv_ctl, 0, sizeof(SurvCtl), stream));
        CU_OK(cudaMemsetAsync(ws->sv_best, 0, size_t(kI8BestM) * 32 * 4, stream));
    }
    static std::mutex mu;
    static std::vector<std::pair<void*, int>> done;
    std::lock_guard<std::mutex> g(mu);
    void* fn = reinterpret_cast<void*>(q1_survivor_kernel(d4));
    bool seen = false;
    for (auto& e : done) seen |= (e.first == fn && e.second == ix->device);
    if (!seen) {
        CU_OK(cudaFuncSetAttribute(q1_survivor_kernel(d4), cudaFuncAttributeMaxDynamicSharedMemorySize, int(ix->smem_optin)));
        done.push_back({fn, ix->device});
    }
    return MVDB_OK;
}

// Survivor mode applies to: one query, 32 < k <= 128 (below that the per-warp selects are as fast: measured), the TMA
// single-query kernel, enough rows to amortise the
// threshold refreshes, and a ring large enough to sort the survivor list in.
static bool sv_eligible(const mvdb_index* ix, const ScanPlan& plan, const ScanParams& p, int g) {
    // (the grid must have more CTAs than k: the threshold is the k-th largest of the per-CTA bests)
    return ix->survivor_tail && g == 1 && plan.tma && plan.q1 && p.k > 32 && p.k <= 128 && p.n >= 16384 && plan.grid > p.k &&
           size_t(p.merge_bytes) >= (size_t(kSurvCap) + 512) * 8 && !p.all_ord;
}

// everything the int8 mode sets up lazily (shadow rows, scratch, shared-memory opt-in): all of it may
// synchronise the device, so a shard group does it before its first launch (prepare_fused_scan)
static int i8_prepare(mvdb_index* ix, mvdb_workspace* ws, uint64_t n, cudaStream_t stream) {
    RC_OK(ensure_shadow8(ix, n));
    RC_OK(ws_overflow_flag(ws));
    if (!ws->i8_ctl) {
        CU_OK(cudaMalloc(&ws->i8_ctl, sizeof(I8Ctl)));
        CU_OK(cudaMalloc(&ws->i8_surv, size_t(kI8SurvCap) * 8));
        CU_OK(cudaMalloc(&ws->i8_best, size_t(kI8BestM) * 32 * 4));
        // zero once; from then on every search leaves the control block and the table clean for the next
        CU_OK(cudaMemsetAsync(ws->i8_ctl, 0, sizeof(I8Ctl), stream));
        CU_OK(cudaMemsetAsync(ws->i8_best, 0, size_t(kI8BestM) * 32 * 4, stream));
    }
    static std::once_flag attr_once[16];
    cudaError_t ae = cudaSuccess;
    std::call_once(attr_once[ix->device & 15], [&] {
        ae = cudaFuncSetAttribute(scan_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ix->smem_optin));
    });
    CU_OK(ae);
    return MVDB_OK;
}

// ONE launch: int8 scan + in-warp exact re-scoring of the candidates + last-CTA sort.  If the survivor list
// overflows (adversarial data: thousands of near-ties) the query is answered by the fp32 scan instead --
// by the caller when it synchronises anyway (host-buffer API: *redo_host), else by a conditional launch
// behind this one (*run_if = device flag; a no-op when the flag is clear).
static int run_i8(mvdb_index* ix, mvdb_workspace* ws, const float* q_dev, int64_t k, const uint32_t* mask_dev, uint32_t n,
                  int normalize_q, int64_t label_offset, float* D_dev, int64_t* I_dev, cudaStream_t stream,
                  const unsigned int** run_if, mvdb_exchange* xch, uint64_t xch_seq) {
    RC_OK(i8_prepare(ix, ws, n, stream));
    I8Params p = {};
    p.x8 = static_cast<const uint8_t*>(ix->mat8.ptr());
    p.x = static_cast<const float*>(ix->mat.ptr());
    p.q = q_dev;
    p.normalize_q = normalize_q;
    p.live = ix->ndead.load(std::memory_order_acquire) ? static_cast<const uint32_t*>(ix->live.ptr()) : nullptr;
    p.mask = mask_dev;
    p.surv = ws->i8_surv;
    p.ovf_host = (tl_host_checks_i8 && !xch) ? ws->i8_ovf_dev : nullptr;
    p.xchg = xch ? xch->dev : nullptr;
    p.xchg_seq = xch_seq;
    p.best = ws->i8_best;
    p.ctl = ws->i8_ctl;
    p.outD = D_dev;
    p.outI = I_dev;
    p.label_offset = label_offset;
    p.n = n;
    p.rec_bytes = ix->rec8;
    p.stage_bytes = uint32_t(align_up(size_t(kI8TileRows) * ix->rec8, 128));
    p.q_off = 1024;
    p.stage_off = uint32_t(1024 + align_up(size_t(ix->ld) * 4, 128));
    p.d = ix->d;
    p.ld4 = ix->ld4;
    p.ld8 = ix->ld8;
    p.k = int(k);
    const int ncw = i8_consumer_warps(ix);
    int stages = int(std::min<size_t>(16, (ix->smem_optin - p.stage_off) / p.stage_bytes));
    stages = stages / ncw * ncw;
    p.stages = stages;
    p.max_norm = std::sqrt(std::max(ix->max_norm2_host.load(std::memory_order_acquire), 0.f)) * 1.0001f;
    const uint32_t tiles = (n + kI8TileRows - 1) / kI8TileRows;
    const int grid = int(std::min<uint32_t>(uint32_t(ix->sm_count), tiles));
    {
        // the first (100 - dyn_tiles) % of every CTA's share is a static round-robin, the rest is claimed from a counter
        const uint32_t per_cta = tiles / uint32_t(grid);
        p.static_iters = ix->dyn_tiles > 0 ? uint32_t(uint64_t(per_cta) * uint32_t(100 - ix->dyn_tiles) / 100u) : 0xFFFFFFFFu;
        p.dyn_tile0 = ix->dyn_tiles > 0 ? p.static_iters * uint32_t(grid) : 0xFFFFFFFFu;
    }
    p.nbest = k <= 32 ? uint32_t(grid) : uint32_t(grid) * uint32_t(ncw);
    if (p.nbest > uint32_t(kI8BestM) * 32) return fail(MVDB_ERR_STATE, "int8 scan: too many consumer warps for the threshold table");
    // the idle ring doubles as the tail's scratch: 4096 survivor keys + the exchange merge's select buffer
    const size_t smem = std::max(size_t(p.stage_off) + size_t(stages) * p.stage_bytes, size_t(p.stage_off) + (size_t(kSelectMax) + 512) * 8);
    if (tl_stage_dep) {   // host path: start under the staging pull
        tl_stage_dep = false;
        p.dep_inputs = 1u;
        CU_OK(launch_dependent(scan_i8_kernel, grid, 32 * (1 + ncw), smem, stream, p));
    } else {
        scan_i8_kernel<<<grid, 32 * (1 + ncw), smem, stream>>>(p);
    }
    LAUNCHED();
    CU_OK(cudaGetLastError());
    *run_if = (tl_host_checks_i8 && !xch) ? nullptr : &ws->i8_ctl->overflow;
    return MVDB_OK;
}

// Core search on device buffers.  Caller holds move_mu shared.
static int run_search(mvdb_index* ix, mvdb_workspace* ws, const float* q_dev, int64_t nq, int64_t k,
                      const uint32_t* mask_dev, uint64_t mask_rows, int normalize_q, int64_t label_offset,
                      float* D_dev, int64_t* I_dev, cudaStream_t stream, mvdb_exchange* xch,
                      const QMaskRef* qmasks) {
    // qmasks: optional per-query admissible bitmasks ([nq], dev == nullptr = unfiltered); used by
    // the coalescer; k <= fused_k_max
    if (nq <= 0) return MVDB_OK;
    if (xch) {
        if (!xch->connected) return fail(MVDB_ERR_STATE, "exchange is not connected");
        if (k > ix->fused_k_max || k > xch->k_max)
            return fail(MVDB_ERR_ARG, "fused exchange supports k <= %d", std::min(ix->fused_k_max, xch->k_max));
    }
    uint64_t n64 = ix->ntotal.load(std::memory_order_acquire);
    if (mask_dev) n64 = std::min<uint64_t>(n64, mask_rows);
    if (n64 > 0xFFFFFFF0ull) return fail(MVDB_ERR_ARG, "more than 2^32 rows per index are not supported");
    const uint32_t n = uint32_t(n64);
    if (n == 0 && xch) {
        // empty shard: still take part in every exchange round
        int64_t done = 0;
        while (done < nq) {
            int64_t rem = nq - done;
            int g = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
            g = std::min(g, xch->nq_max);
            xchg_empty_kernel<<<1, 128, size_t(4) * select_cap(int(k)) * 8, stream>>>(xch->dev, ++xch->seq, g, int(k),
                                                                                      D_dev + done * k, I_dev + done * k);
            LAUNCHED();
            done += g;
        }
        CU_OK(cudaGetLastError());
        return MVDB_OK;
    }
    if (n == 0) {
        int64_t total = nq * k;
        fill_empty_results_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(D_dev, I_dev, total);
        LAUNCHED();
        CU_OK(cudaGetLastError());
        return MVDB_OK;
    }
    bool any_qmask = false;
    if (qmasks)
        for (int64_t i = 0; i < nq; i++) any_qmask |= qmasks[i].dev != nullptr;
    if (any_qmask && k > ix->fused_k_max) return fail(MVDB_ERR_ARG, "per-query masks need k <= %d", ix->fused_k_max);
    // Scan or tensor cores?  A measured cost model (B200): the fp32 scan serves 1/2/4/8 queries
    // per pass at 1.0/1.08/1.2/2.3 x the stream time of the matrix plus ~40 us per launch; the
    // batched path costs ~330 us of fixed work (levels, threshold refreshes, re-scoring) plus one
    // pass over the half-size bf16 shadow plus the GEMM itself.  `batch_min_nq` is only a floor.
    bool use_tc = false;
    {
        const double bytes32 = double(n) * double(ix->ld) * 4.0;
        double t_scan = 0.0;
        for (int64_t rem = nq; rem > 0;) {
            const int g = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
            const double f = g == 8 ? 2.3 : g == 4 ? 1.2 : g == 2 ? 1.08 : 1.0;
            t_scan += 40e-6 + f * bytes32 / 6.4e12;
            rem -= g;
        }
        // one pass over the bf16 shadow at ~900 TFLOP/s effective -- or, in tf32 mode, over the
        // fp32 matrix itself at half the tensor rate
        const bool tf32 = ix->batch_mode == 3;
        const double t_tc = 330e-6 + 1.07 * (bytes32 * (tf32 ? 1.0 : 0.5)) / 6.0e12 +
                            2.0 * double(nq) * double(n) * double(ix->d) / (tf32 ? 450e12 : 900e12);
        use_tc = nq >= ix->batch_min_nq && (ix->batch_cost_model ? (nq >= 2 && t_tc < t_scan) : true);
    }
    if (!xch && !tl_force_scan && ix->batch_mode != 0 && use_tc && k <= 128 &&
        tensor_map_encoder() != nullptr)
        return run_batched(ix, ws, q_dev, nq, k, mask_dev, n, normalize_q, label_offset, D_dev, I_dev, stream,
                           ix->batch_mode, nullptr, any_qmask ? qmasks : nullptr);
    RC_OK(ws_scratch(ws));
    // opt-in int8 shadow mode: the whole query is answered by the shadow pass + exact re-scoring; the fp32
    // scan below is still enqueued, as a conditional launch that only runs if a candidate list overflowed
    const unsigned int* run_if = nullptr;
    uint64_t i8_seq = 0;   // sharded search: the int8 launch and its conditional fp32 fallback are ONE exchange round
    if (!tl_force_scan && !any_qmask && i8_eligible(ix, nq, k, n)) {
        if (xch) i8_seq = ++xch->seq;
        RC_OK(run_i8(ix, ws, q_dev, k, mask_dev, n, normalize_q, label_offset, D_dev, I_dev, stream, &run_if, xch, i8_seq));
        if (!run_if) return MVDB_OK;   // the caller synchronises, checks the overflow word and re-runs on the fp32 scan if needed
    }
    ScanParams p = {};
    p.run_if = run_if;
    p.x = static_cast<const float*>(ix->mat.ptr());
    p.live = ix->ndead.load(std::memory_order_acquire) ? static_cast<const uint32_t*>(ix->live.ptr()) : nullptr;
    p.mask = mask_dev;
    p.ticket = ws->ticket;
    p.label_offset = label_offset;
    p.n = n;
    p.d = ix->d;
    p.ld4 = ix->ld4;
    p.normalize_q = normalize_q;
    p.trace = ix->trace_dev;
    p.pin_tiles = uint32_t((uint64_t(ix->l2_pin_mb) << 20) / (uint64_t(kRowsPerTile) * ix->ld * 4));

    if (k <= ix->fused_k_max) {
        p.k = int(k);
        int64_t done = 0;
        while (done < nq) {
            int64_t rem = nq - done;
            int g = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
            if (xch) {
                while (g > xch->nq_max) g >>= 1;
                p.xchg = xch->dev;
                p.xchg_seq = i8_seq ? i8_seq : ++xch->seq;
            }
            p.mask = mask_dev;
            p.n = n;
            p.has_qmask = 0;
            for (int i = 0; i < 8; i++) p.qmask[i] = nullptr;
            if (any_qmask) {
                if (g == 1) {
                    if (qmasks[done].dev) {   // the single-query kernel takes it as the common mask
                        p.mask = qmasks[done].dev;
                        p.n = uint32_t(std::min<uint64_t>(n, uint64_t(qmasks[done].words) * 32));
                        if (p.n == 0) {
                            fill_empty_results_kernel<<<unsigned((k + 255) / 256), 256, 0, stream>>>(D_dev + done * k, I_dev + done * k, k);
                            LAUNCHED();
                            done += 1;
                            continue;
                        }
                    }
                } else {
                    for (int i = 0; i < g; i++) {
                        p.qmask[i] = qmasks[done + i].dev;
                        p.qmask_bytes[i] = qmasks[done + i].words * 4u;
                        p.has_qmask |= p.qmask[i] != nullptr;
                    }
                }
            }
            ScanPlan plan;
            RC_OK(plan_scan(ix, p, g, &plan));
            RC_OK(grow_dev(&ws->partials, &ws->partials_cap, size_t(g) * plan.grid * p.k));
            p.partials = ws->partials;
            p.q = q_dev + done * ix->d;
            p.outD = D_dev + done * k;
            p.outI = I_dev + done * k;
            p.all_ord = nullptr;
            p.tile_ctr = nullptr;
            if (ix->dyn_tiles > 0 && plan.tma) {
                // the first (100 - dyn_tiles)% of every CTA's share is a static round-robin, the
                // rest is claimed from a counter in chunks aligned to one 32-row mask word
                const uint32_t tiles = (p.n + kRowsPerTile - 1) / kRowsPerTile;
                uint32_t st = uint32_t(uint64_t(tiles / uint32_t(plan.grid)) * uint32_t(100 - ix->dyn_tiles) / 100u);
                if (plan.grid & 3) st &= ~3u;
                p.static_iters = st;
                p.dyn_tile0 = st * uint32_t(plan.grid);
                // two counters, alternating per launch: with "pdl" the next scan claims tiles while this one is still finishing
                p.tile_ctr = ws->ticket + 32 + 32 * (ws->launch_seq++ & 1u);
            }
            p.pdl_early = 0;
            const unsigned int* launch_if = run_if;
            if (!run_if && !tl_force_scan && sv_eligible(ix, plan, p, g)) {
                // 32 < k <= 128: shared threshold + global survivor list, the last CTA sorts; the classic kernel is
                // only its overflow fallback (conditional launch, or the host caller re-runs the query)
                const int d4 = (ix->ld4 + 31) / 32;
                RC_OK(sv_prepare(ix, ws, d4, stream));
                ScanParams sp = p;
                sp.surv = ws->sv_surv;
                sp.sctl = ws->sv_ctl;
                sp.best = ws->sv_best;
                sp.surv_cap = kSurvCap;
                sp.nbest = uint32_t(plan.grid);   // one slot per CTA
                sp.ovf_host = (tl_host_checks_i8 && !xch) ? ws->i8_ovf_dev : nullptr;
                if (sp.nbest <= uint32_t(kI8BestM) * 32) {
                    if (tl_stage_dep) {
                        tl_stage_dep = false;
                        sp.dep_inputs = 1u;
                        CU_OK(launch_dependent(q1_survivor_kernel(d4), plan.grid, plan.threads, plan.smem, stream, sp));
                    } else {
                        q1_survivor_kernel(d4)<<<plan.grid, plan.threads, plan.smem, stream>>>(sp);
                    }
                    LAUNCHED();
                    CU_OK(cudaGetLastError());
                    if (tl_host_checks_i8 && !xch) {   // the caller checks the pinned flag after its synchronise
                        done += g;
                        continue;
                    }
                    launch_if = &ws->sv_ctl->overflow;
                    if (p.tile_ctr) p.tile_ctr = ws->ticket + 32 + 32 * (ws->launch_seq++ & 1u);   // the fallback's own counter slot
                }
            }
            p.run_if = launch_if;
            if (ix->pdl && tl_allow_pdl && !launch_if) {
                // Entry trigger only when this launch AND the previous scan of this workspace fill every SM
                // with one CTA each: then a CTA of this grid can only start where the previous grid has left,
                // all of them have started only once the previous grid is complete, and the grid after this
                // one (which waits for all of ours to start) cannot overlap the previous one.
                const bool big = 2 * plan.smem + 2048 > size_t(ix->smem_per_sm) && plan.grid >= ix->sm_count;
                p.pdl_early = (big && ws->prev_scan_big) ? 1u : 0u;
                ws->prev_scan_big = big;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(unsigned(plan.grid));
                cfg.blockDim = dim3(unsigned(plan.threads));
                cfg.dynamicSmemBytes = plan.smem;
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                CU_OK(cudaLaunchKernelEx(&cfg, plan.fn, p));
            } else if (tl_stage_dep && plan.q1) {
                tl_stage_dep = false;
                ws->prev_scan_big = false;
                p.dep_inputs = 1u;
                CU_OK(launch_dependent(plan.fn, plan.grid, plan.threads, plan.smem, stream, p));
                p.dep_inputs = 0u;
            } else {
                tl_stage_dep = false;   // any other kernel runs in plain stream order behind the pull
                ws->prev_scan_big = false;
                plan.fn<<<plan.grid, plan.threads, plan.smem, stream>>>(p);
            }
            LAUNCHED();
            CU_OK(cudaGetLastError());
            done += g;
        }
        return MVDB_OK;
    }

    // ---- large k: materialise score images, radix-select, sort -------------
    if (!ws->radix) CU_OK(cudaMalloc(&ws->radix, sizeof(RadixState)));
    RC_OK(grow_dev(&ws->all_ord, &ws->all_ord_cap, size_t(n)));
    const uint64_t kk = std::min<uint64_t>(uint64_t(k), n);
    uint64_t npad = 1;
    while (npad < kk) npad <<= 1;
    RC_OK(grow_dev(&ws->keys, &ws->keys_cap, size_t(npad)));
    p.k = 1;
    const int hist_grid = int(std::min<uint64_t>(uint64_t(ix->sm_count) * 4, (uint64_t(n) + 511) / 512));
    ws->prev_scan_big = false;
    // host-buffer callers (they synchronise and can re-run a query): two 12-bit histogram passes, collect, one sort
    // -- 5 launches instead of 22; the radix select below is the fallback when too many rows tie around the k-th
    const bool fsel = tl_host_checks_i8 && !tl_force_scan && !xch && ix->large_k_fast && kk <= 8192;
    if (fsel) {
        RC_OK(ws_overflow_flag(ws));
        if (!ws->fsel) {
            CU_OK(cudaMalloc(&ws->fsel, sizeof(FselState)));
            CU_OK(cudaMemsetAsync(ws->fsel, 0, sizeof(FselState), stream));
        }
        RC_OK(grow_dev(&ws->keys, &ws->keys_cap, size_t(kFselCap)));
        CU_OK(cudaFuncSetAttribute(fsel_sort_results_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFselCap * 8)));
    }
    for (int64_t qi = 0; qi < nq; qi++) {
        ScanPlan plan;
        RC_OK(plan_scan(ix, p, 1, &plan));
        p.q = q_dev + qi * ix->d;
        p.all_ord = ws->all_ord;
        p.partials = nullptr;
        p.outD = nullptr;
        p.outI = nullptr;
        plan.fn<<<plan.grid, plan.threads, plan.smem, stream>>>(p);
        LAUNCHED();
        if (fsel) {
            fsel_hist_kernel<<<hist_grid, 512, 0, stream>>>(ws->all_ord, n, ws->fsel, uint32_t(kk), 0);
            LAUNCHED();
            fsel_hist_kernel<<<hist_grid, 512, 0, stream>>>(ws->all_ord, n, ws->fsel, uint32_t(kk), 1);
            LAUNCHED();
            fsel_collect_kernel<<<hist_grid, 512, 0, stream>>>(ws->all_ord, n, ws->fsel, uint32_t(kk), ws->keys);
            LAUNCHED();
            fsel_sort_results_kernel<<<1, 1024, size_t(kFselCap) * 8, stream>>>(ws->keys, ws->fsel, k, D_dev + qi * k, I_dev + qi * k,
                                                                              label_offset, ws->i8_ovf_dev);
            LAUNCHED();
            CU_OK(cudaGetLastError());
            continue;
        }
        radix_init_kernel<<<1, 256, 0, stream>>>(ws->radix, kk);
        LAUNCHED();
        for (int shift = 56; shift >= 0; shift -= 8) {
            radix_hist_kernel<<<hist_grid, 512, 0, stream>>>(ws->all_ord, n, ws->radix, shift);
            radix_pick_kernel<<<1, 256, 0, stream>>>(ws->radix, shift);
            LAUNCHED();
            LAUNCHED();
        }
        radix_collect_kernel<<<hist_grid, 512, 0, stream>>>(ws->all_ord, n, ws->radix, ws->keys, uint32_t(kk));
        LAUNCHED();
        pad_keys_kernel<<<unsigned((npad + 255) / 256), 256, 0, stream>>>(ws->keys, ws->radix, uint32_t(npad));
        LAUNCHED();
        if (npad <= 16384) {
            cudaFuncSetAttribute(sort_keys_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
            sort_keys_smem_kernel<<<1, 1024, npad * 8, stream>>>(ws->keys, uint32_t(npad));
            LAUNCHED();
        } else {
            for (uint64_t size = 2; size <= npad; size <<= 1)
                for (uint64_t stride = size >> 1; stride > 0; stride >>= 1) {
                    sort_keys_global_step_kernel<<<unsigned((npad / 2 + 255) / 256), 256, 0, stream>>>(ws->keys, npad, size, stride);
                    LAUNCHED();
                }
        }
        keys_to_results_kernel<<<unsigned((k + 255) / 256), 256, 0, stream>>>(ws->keys, uint32_t(kk), k, D_dev + qi * k,
                                                                               I_dev + qi * k, label_offset);
        LAUNCHED();
        CU_OK(cudaGetLastError());
    }
    return MVDB_OK;
}

// ---------------------------------------------------------------------------
// workspace handling
// ---------------------------------------------------------------------------
static void ws_release(mvdb_workspace* ws) {   // device / pinned memory and the stream; the struct stays
    cudaFree(ws->partials);
    cudaFree(ws->ticket);
    cudaFree(ws->all_ord);
    cudaFree(ws->radix);
    cudaFree(ws->fsel);
    cudaFree(ws->keys);
    cudaFree(ws->b_qn);
    cudaFree(ws->b_qnorm);
    cudaFree(ws->b_q16);
    cudaFree(ws->b_thr);
    cudaFree(ws->b_cnt);
    cudaFree(ws->b_ovf);
    cudaFree(ws->b_cand);
    cudaFree(ws->b_qmptr);
    cudaFree(ws->b_qmwords);
    cudaFree(ws->b_lmask);
    cudaFree(ws->b_lqmask);
    cudaFree(ws->b_lqptr);
    cudaFree(ws->b_lqwords);
    cudaFree(ws->sv_surv);
    cudaFree(ws->sv_best);
    cudaFree(ws->sv_ctl);
    cudaFree(ws->i8_surv);
    if (ws->i8_ovf_pin) cudaFreeHost(ws->i8_ovf_pin);
    cudaFree(ws->i8_best);
    cudaFree(ws->i8_ctl);
    cudaFree(ws->q_dev);
    cudaFree(ws->mask_dev);
    cudaFree(ws->I_dev);
    cudaFreeHost(ws->q_pin);
    cudaFreeHost(ws->mask_pin);
    cudaFreeHost(ws->I_pin);
    if (ws->stream) cudaStreamDestroy(ws->stream);
    *ws = mvdb_workspace();
}
static void ws_free(mvdb_workspace* ws) {
    if (!ws) return;
    ws_release(ws);
    delete ws;
}

static int ws_new(mvdb_index* ix, mvdb_workspace** out) {
    mvdb_workspace* ws = new mvdb_workspace();
    ws->ix = ix;
    cudaError_t e = cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ws;
        return fail(MVDB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    *out = ws;
    return MVDB_OK;
}

static int pool_acquire(mvdb_index* ix, mvdb_workspace** out) {
    std::unique_lock<std::mutex> lk(ix->pool_mu);
    for (;;) {
        if (!ix->pool_free.empty()) {
            *out = ix->pool_free.back();
            ix->pool_free.pop_back();
            return MVDB_OK;
        }
        if (ix->pool_created < mvdb_index::kPoolMax) {
            ix->pool_created++;
            lk.unlock();
            int rc = ws_new(ix, out);
            if (rc != MVDB_OK) {
                lk.lock();
                ix->pool_created--;
                return rc;
            }
            (*out)->pooled = true;
            return MVDB_OK;
        }
        ix->pool_cv.wait(lk);
    }
}
static void pool_release(mvdb_index* ix, mvdb_workspace* ws) {
    {
        std::lock_guard<std::mutex> g(ix->pool_mu);
        ix->pool_free.push_back(ws);
    }
    ix->pool_cv.notify_one();
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int mvdb_abi_version(void) { return MVDB_ABI_VERSION; }
const char* mvdb_last_error(void) { return g_err.c_str(); }
uint64_t mvdb_launch_count(void) { return g_launches.load(); }

int mvdb_device_count(int* count) {
    if (!count) return fail(MVDB_ERR_ARG, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
    }
    *count = n;
    return MVDB_OK;
}

int mvdb_index_create(int d, int device, uint64_t capacity_hint, mvdb_index** out) {
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    *out = nullptr;
    if (d <= 0 || d > (1 << 20)) return fail(MVDB_ERR_ARG, "dimension %d out of range", d);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(MVDB_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(MVDB_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MVDB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    CU_OK(cudaFree(0));
    cudaDeviceProp prop;
    CU_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(MVDB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    mvdb_index* ix = new mvdb_index();
    ix->d = d;
    ix->device = device;
    ix->ld = int64_t(align_up(size_t(d), 4));
    ix->ld4 = int(ix->ld / 4);
    ix->sm_count = prop.multiProcessorCount;
    ix->smem_optin = prop.sharedMemPerBlockOptin;
    ix->smem_per_sm = prop.sharedMemPerMultiprocessor;
    if (const char* sl = getenv("MVDB_SMEM_SLACK")) ix->smem_optin -= size_t(atoi(sl));
    size_t free_b = 0, total_b = 0;
    CU_OK(cudaMemGetInfo(&free_b, &total_b));
    size_t reserve = capacity_hint ? size_t(capacity_hint) * ix->ld * 4 : total_b;
    reserve = std::max(reserve, size_t(1) << 21);
    int rc = ix->mat.init(device, reserve);
    if (rc == MVDB_OK) rc = ix->live.init(device, std::max<size_t>(reserve / (size_t(ix->ld) * 4) / 8 + 4096, size_t(1) << 21));
    ix->ld8 = int(align_up(size_t(d), 16));
    ix->rec8 = uint32_t(ix->ld8 + kI8MetaBytes);
    if (rc == MVDB_OK) rc = ix->mat8.init(device, std::max<size_t>(reserve / (size_t(ix->ld) * 4) * ix->rec8, size_t(1) << 21));
    ix->ld16 = int64_t(align_up(size_t(d), 8));
    if (rc == MVDB_OK) rc = ix->mat16.init(device, std::max<size_t>(reserve / (size_t(ix->ld) * 4) * ix->ld16 * 2, size_t(1) << 21));
    cudaError_t e = cudaStreamCreateWithFlags(&ix->mut_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ix->filter_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ix->max_norm2_bits, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ix->max_norm2_bits, 0, sizeof(int));
    if (rc != MVDB_OK || e != cudaSuccess) {
        ix->mat.destroy();
        ix->live.destroy();
        ix->mat16.destroy();
        ix->mat8.destroy();
        delete ix;
        return rc != MVDB_OK ? rc : fail(MVDB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    *out = ix;
    return MVDB_OK;
}

int mvdb_index_destroy(mvdb_index* ix) {
    if (!ix) return MVDB_OK;
    if (ix->group) return fail(MVDB_ERR_STATE, "this index belongs to a shard group: destroy the group first");
    DeviceGuard guard(ix->device);
    cudaDeviceSynchronize();
    {
        // orphan everything that points back at us: their memory goes now, their structs when their owner destroys them
        std::lock_guard<std::mutex> g(ix->child_mu);
        for (mvdb_mask* m : ix->child_masks) {
            cudaFree(m->dev);
            if (m->ready) cudaEventDestroy(m->ready);
            m->dev = nullptr;
            m->ready = nullptr;
            m->ix = nullptr;
        }
        for (mvdb_column* c : ix->child_columns) {
            cudaFree(c->vals);
            cudaFree(c->has);
            c->vals = nullptr;
            c->has = nullptr;
            c->ix = nullptr;
        }
        for (mvdb_workspace* ws : ix->child_workspaces) ws_release(ws);   // resets the struct: ws->ix becomes null
        ix->child_masks.clear();
        ix->child_columns.clear();
        ix->child_workspaces.clear();
    }
    for (auto* ws : ix->pool_free) ws_free(ws);
    ix->pool_free.clear();
    ix->mat.destroy();
    ix->live.destroy();
    ix->mat16.destroy();
    ix->mat8.destroy();
    cudaFree(ix->max_norm2_bits);
    cudaFree(ix->count_dev);
    cudaFree(ix->stage_dev[0]);
    cudaFree(ix->stage_dev[1]);
    cudaFree(ix->rows_dev);
    if (ix->mut_stream) cudaStreamDestroy(ix->mut_stream);
    if (ix->filter_stream) cudaStreamDestroy(ix->filter_stream);
    delete ix;
    return MVDB_OK;
}

int mvdb_index_reset(mvdb_index* ix) {
    ENTER(ix);
    std::unique_lock<std::shared_mutex> mv(ix->move_mu);
    std::lock_guard<std::mutex> g(ix->mut_mu);
    size_t words = (ix->ntotal.load() + 31) / 32;
    if (words) CU_OK(cudaMemsetAsync(ix->live.ptr(), 0, words * 4, ix->mut_stream));
    CU_OK(cudaStreamSynchronize(ix->mut_stream));
    ix->live_host.clear();
    ix->ntotal.store(0);
    ix->ndead.store(0);
    {
        std::lock_guard<std::mutex> sg(ix->shadow_mu);
        ix->shadow_rows = 0;
    }
    {
        std::lock_guard<std::mutex> sg(ix->shadow8_mu);
        ix->shadow8_rows = 0;
    }
    return MVDB_OK;
}

int mvdb_index_set_option(mvdb_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return fail(MVDB_ERR_ARG, "null argument");
    std::string s(name);
    if (s == "scan_variant") {
        if (value < 0 || value > 2) return fail(MVDB_ERR_ARG, "scan_variant must be 0..2");
        ix->scan_variant = int(value);
    } else if (s == "fused_k_max") {
        if (value < 0 || value > 1024) return fail(MVDB_ERR_ARG, "fused_k_max must be 0..1024");
        ix->fused_k_max = int(value);
    } else if (s == "grid_ctas") {
        if (value < 0 || value > 65535) return fail(MVDB_ERR_ARG, "grid_ctas out of range");
        ix->grid_ctas = int(value);
    } else if (s == "coalesce") {
        if (value < 0 || value > 1) return fail(MVDB_ERR_ARG, "coalesce must be 0 or 1");
        ix->coalesce = int(value);
    } else if (s == "coalesce_wait_pct") {
        if (value < 0 || value > 50) return fail(MVDB_ERR_ARG, "coalesce_wait_pct must be 0..50");
        ix->co_wait_pct = int(value);
    } else if (s == "coalesce_max") {
        if (value < 1 || value > 1024) return fail(MVDB_ERR_ARG, "coalesce_max must be 1..1024");
        ix->coalesce_max = int(value);
    } else if (s == "batch_mode") {
        if (value < 0 || value > 3) return fail(MVDB_ERR_ARG, "batch_mode must be 0 (off), 1 (exact), 2 (bf16) or 3 (tf32)");
        ix->batch_mode = int(value);
    } else if (s == "gemm_short_a") {
        ix->gemm_short_a = value != 0;
    } else if (s == "gemm_debug") {
        ix->gemm_debug = int(value);
    } else if (s == "gemm_prof") {
        if (value && !ix->gemm_prof_dev) {
            CU_OK(cudaMalloc(&ix->gemm_prof_dev, 256 * 8 * 8));
            CU_OK(cudaMemset(ix->gemm_prof_dev, 0, 256 * 8 * 8));
        } else if (!value && ix->gemm_prof_dev) {
            cudaFree(ix->gemm_prof_dev);
            ix->gemm_prof_dev = nullptr;
        }
    } else if (s == "trace") {
        DeviceGuard guard(ix->device);
        if (value && !ix->trace_dev) {
            CU_OK(cudaMalloc(&ix->trace_dev, 16 * 8));
            CU_OK(cudaMemset(ix->trace_dev, 0, 16 * 8));
        } else if (!value && ix->trace_dev) {
            cudaFree(ix->trace_dev);
            ix->trace_dev = nullptr;
        }
    } else if (s == "coalesce_leaders") {
        if (value < 0 || value > 2) return fail(MVDB_ERR_ARG, "coalesce_leaders must be 0 (auto), 1 or 2");
        ix->co_max_leaders = int(value);
    } else if (s == "large_k_fast") {
        ix->large_k_fast = value != 0;
    } else if (s == "host_path") {
        if (value < 0 || value > 3) return fail(MVDB_ERR_ARG, "host_path is a bit set 0..3");
        ix->host_path = int(value);
    } else if (s == "pdl") {
        ix->pdl = value != 0;
    } else if (s == "survivor_tail") {
        ix->survivor_tail = value != 0;
    } else if (s == "scan_shadow") {
        if (value < 0 || value > 1) return fail(MVDB_ERR_ARG, "scan_shadow must be 0 (fp32 scan) or 1 (int8 shadow + exact fp32 re-scoring)");
        ix->scan_shadow = int(value);
    } else if (s == "dyn_tiles") {
        if (value < 0 || value > 100) return fail(MVDB_ERR_ARG, "dyn_tiles is a percentage, 0..100");
        ix->dyn_tiles = int(value);
    } else if (s == "l2_pin_mb") {
        if (value < 0 || value > 120) return fail(MVDB_ERR_ARG, "l2_pin_mb must be 0..120");
        ix->l2_pin_mb = int(value);
    } else if (s == "gemm_variant") {
        if (value < 0 || value > 3) return fail(MVDB_ERR_ARG, "gemm_variant must be 0..3");
        ix->gemm_variant = int(value);
    } else if (s == "gemm_l2_hint") {
        if (value < 0 || value > 2) return fail(MVDB_ERR_ARG, "gemm_l2_hint must be 0..2");
        ix->gemm_l2_hint = int(value);
    } else if (s == "batch_cost_model") {
        ix->batch_cost_model = value != 0;
    } else if (s == "batch_min_nq") {
        if (value < 1) return fail(MVDB_ERR_ARG, "batch_min_nq must be >= 1");
        ix->batch_min_nq = int(value);
    } else if (s == "consumer_warps") {
        if (value < 0 || value > 8) return fail(MVDB_ERR_ARG, "consumer_warps must be 0..8");
        ix->consumer_warps = int(value);
    } else {
        return fail(MVDB_ERR_ARG, "unknown option '%s'", name);
    }
    return MVDB_OK;
}

// shared tail of the three add flavours: capacity, kernels, bookkeeping
static int add_common(mvdb_index* ix, const float* x_host, const float* x_dev, bool synth, uint64_t seed,
                      int64_t synth_row0, int dist, uint64_t n, int normalize, int64_t* first_row) {
    std::unique_lock<std::shared_mutex> mv(ix->move_mu, std::defer_lock);
    if (!ix->mat.stable()) mv.lock();  // cudaMalloc fallback may move the matrix
    std::lock_guard<std::mutex> g(ix->mut_mu);
    const uint64_t n0 = ix->ntotal.load();
    if (first_row) *first_row = int64_t(n0);
    if (n == 0) return MVDB_OK;
    if (n0 + n > 0xFFFFFFF0ull) return fail(MVDB_ERR_ARG, "more than 2^32 rows per index are not supported");
    cudaStream_t st = ix->mut_stream;
    RC_OK(ix->mat.ensure(size_t(n0 + n) * ix->ld * 4, st));
    RC_OK(ix->live.ensure(align_up((n0 + n + 31) / 32 * 4, 256), st));
    float* base = static_cast<float*>(ix->mat.ptr()) + n0 * ix->ld;
    const int threads = 256;
    auto grid_for = [&](uint64_t rows) { return unsigned(std::min<uint64_t>((rows * 32 + threads - 1) / threads, uint64_t(ix->sm_count) * 16)); };
    if (synth) {
        append_rows_kernel<true><<<grid_for(n), threads, 0, st>>>(nullptr, base, n, ix->d, ix->ld, normalize, seed,
                                                                  uint64_t(synth_row0), dist, ix->max_norm2_bits);
        LAUNCHED();
    } else if (x_dev) {
        append_rows_kernel<false><<<grid_for(n), threads, 0, st>>>(x_dev, base, n, ix->d, ix->ld, normalize, 0, 0, 0,
                                                                   ix->max_norm2_bits);
        LAUNCHED();
    } else {
        // host rows: stream through two device staging buffers so the H2D copy
        // of chunk i+1 overlaps the normalise/append kernel of chunk i
        const size_t row_bytes = size_t(ix->d) * 4;
        const uint64_t chunk_rows = std::max<uint64_t>(1, (size_t(64) << 20) / row_bytes);
        uint64_t done = 0;
        int b = 0;
        while (done < n) {
            uint64_t m = std::min(chunk_rows, n - done);
            RC_OK(grow_dev(&ix->stage_dev[b], &ix->stage_cap[b], size_t(m) * ix->d));
            CU_OK(cudaMemcpyAsync(ix->stage_dev[b], x_host + done * ix->d, m * row_bytes, cudaMemcpyHostToDevice, st));
            append_rows_kernel<false><<<grid_for(m), threads, 0, st>>>(ix->stage_dev[b], base + done * ix->ld, m, ix->d,
                                                                       ix->ld, normalize, 0, 0, 0, ix->max_norm2_bits);
            LAUNCHED();
            done += m;
            b ^= 1;
        }
    }
    {
        uint64_t words = ((n0 + n + 31) >> 5) - (n0 >> 5);
        set_live_range_kernel<<<unsigned((words + 255) / 256), 256, 0, st>>>(static_cast<uint32_t*>(ix->live.ptr()), n0, n);
        LAUNCHED();
    }
    CU_OK(cudaGetLastError());
    // int8 shadow mode: keep an up-to-date shadow up to date HERE (same stream, same synchronise) -- otherwise
    // every search that follows an insert would have to extend it first and wait on this busy stream
    bool shadow8_extended = false;
    if (ix->scan_shadow) {
        std::lock_guard<std::mutex> sg(ix->shadow8_mu);
        if (ix->shadow8_rows > 0 && ix->shadow8_rows <= n0) {
            // from wherever the shadow ends (adds that raced a search's own extension leave a gap) up to the new end
            const uint64_t r0 = ix->shadow8_rows, m = n0 + n - r0;
            RC_OK(ix->mat8.ensure(size_t(n0 + n) * ix->rec8, st));
            unsigned grid8 = unsigned(std::min<uint64_t>((m * 32 + 255) / 256, uint64_t(ix->sm_count) * 16));
            to_i8_rows_kernel<<<grid8, 256, 0, st>>>(static_cast<const float*>(ix->mat.ptr()) + r0 * ix->ld,
                                                     static_cast<uint8_t*>(ix->mat8.ptr()) + r0 * ix->rec8, m, ix->ld4, ix->ld8, ix->rec8);
            LAUNCHED();
            CU_OK(cudaGetLastError());
            shadow8_extended = true;
        }
    }
    // same for the bf16 shadow of the tensor-core batch path: under insert churn every coalesced batch used to
    // extend it first and wait for this (busy) stream -- more than doubling the time of a pass
    bool shadow16_extended = false;
    {
        std::lock_guard<std::mutex> sg(ix->shadow_mu);
        if (ix->shadow_rows > 0 && ix->shadow_rows <= n0) {
            const uint64_t r0 = ix->shadow_rows, m = n0 + n - r0;
            RC_OK(ix->mat16.ensure(size_t(n0 + n) * ix->ld16 * 2, st));
            unsigned grid16 = unsigned(std::min<uint64_t>((m * 32 + 255) / 256, uint64_t(ix->sm_count) * 16));
            to_bf16_rows_kernel<<<grid16, 256, 0, st>>>(static_cast<const float*>(ix->mat.ptr()) + r0 * ix->ld,
                                                        static_cast<__nv_bfloat16*>(ix->mat16.ptr()) + r0 * ix->ld16, m, ix->d,
                                                        ix->ld, ix->ld16);
            LAUNCHED();
            CU_OK(cudaGetLastError());
            shadow16_extended = true;
        }
    }
    {
        int bits = 0;
        CU_OK(cudaMemcpyAsync(&bits, ix->max_norm2_bits, sizeof bits, cudaMemcpyDeviceToHost, st));
        CU_OK(cudaStreamSynchronize(st));
        float f;
        memcpy(&f, &bits, 4);
        ix->max_norm2_host.store(f, std::memory_order_release);
    }
    ix->live_host.resize((n0 + n + 31) / 32, 0u);
    for (uint64_t r = n0; r < n0 + n;) {
        if ((r & 31) == 0 && r + 32 <= n0 + n) {
            ix->live_host[r >> 5] = 0xFFFFFFFFu;
            r += 32;
        } else {
            ix->live_host[r >> 5] |= 1u << (r & 31);
            r++;
        }
    }
    if (shadow16_extended) {
        std::lock_guard<std::mutex> sg(ix->shadow_mu);
        if (ix->shadow_rows > 0 && ix->shadow_rows <= n0 + n) ix->shadow_rows = n0 + n;
    }
    if (shadow8_extended) {
        std::lock_guard<std::mutex> sg(ix->shadow8_mu);
        if (ix->shadow8_rows > 0 && ix->shadow8_rows <= n0 + n) ix->shadow8_rows = n0 + n;
    }
    ix->ntotal.store(n0 + n, std::memory_order_release);
    return MVDB_OK;
}

int mvdb_index_add(mvdb_index* ix, const float* x, uint64_t n, int normalize, int64_t* first_row) {
    ENTER(ix);
    if (n && !x) return fail(MVDB_ERR_ARG, "null rows");
    return add_common(ix, x, nullptr, false, 0, 0, 0, n, normalize, first_row);
}
int mvdb_index_add_device(mvdb_index* ix, const float* x_dev, uint64_t n, int normalize, int64_t* first_row) {
    ENTER(ix);
    if (n && !x_dev) return fail(MVDB_ERR_ARG, "null rows");
    return add_common(ix, nullptr, x_dev, false, 0, 0, 0, n, normalize, first_row);
}
int mvdb_index_add_synthetic(mvdb_index* ix, uint64_t seed, int64_t row0, uint64_t n, int dist, int normalize,
                             int64_t* first_row) {
    ENTER(ix);
    if (dist < 0 || dist > 1 || row0 < 0) return fail(MVDB_ERR_ARG, "bad synthetic parameters");
    return add_common(ix, nullptr, nullptr, true, seed, row0, dist, n, normalize, first_row);
}

int mvdb_index_remove_rows(mvdb_index* ix, const int64_t* rows, uint64_t n) {
    ENTER(ix);
    if (n && !rows) return fail(MVDB_ERR_ARG, "null rows");
    std::lock_guard<std::mutex> g(ix->mut_mu);
    const uint64_t nt = ix->ntotal.load();
    // validate everything before changing anything
    std::vector<int64_t> sorted(rows, rows + n);
    std::sort(sorted.begin(), sorted.end());
    for (uint64_t i = 0; i < n; i++) {
        int64_t r = sorted[i];
        if (r < 0 || uint64_t(r) >= nt) return fail(MVDB_ERR_ARG, "row %lld out of range", (long long)r);
        if (i && sorted[i - 1] == r) return fail(MVDB_ERR_ARG, "row %lld listed twice", (long long)r);
        if (!((ix->live_host[r >> 5] >> (r & 31)) & 1u)) return fail(MVDB_ERR_ARG, "row %lld already deleted", (long long)r);
    }
    if (n == 0) return MVDB_OK;
    cudaStream_t st = ix->mut_stream;
    RC_OK(grow_dev(&ix->rows_dev, &ix->rows_cap, size_t(n)));
    CU_OK(cudaMemcpyAsync(ix->rows_dev, rows, n * 8, cudaMemcpyHostToDevice, st));
    // publish "tombstones exist" before the bits flip so new searches read the bitmask
    ix->ndead.fetch_add(n, std::memory_order_release);
    clear_live_rows_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(static_cast<uint32_t*>(ix->live.ptr()), ix->rows_dev, n);
    LAUNCHED();
    CU_OK(cudaGetLastError());
    CU_OK(cudaStreamSynchronize(st));
    for (uint64_t i = 0; i < n; i++) ix->live_host[rows[i] >> 5] &= ~(1u << (rows[i] & 31));
    return MVDB_OK;
}

int mvdb_index_compact(mvdb_index* ix, int64_t* ntotal_out) {
    ENTER(ix);
    std::unique_lock<std::shared_mutex> mv(ix->move_mu);  // waits for running searches
    std::lock_guard<std::mutex> g(ix->mut_mu);
    const uint64_t nt = ix->ntotal.load();
    if (ix->ndead.load() == 0) {
        if (ntotal_out) *ntotal_out = int64_t(nt);
        return MVDB_OK;
    }
    std::vector<uint32_t> src;
    src.reserve(nt);
    for (uint64_t r = 0; r < nt; r++)
        if ((ix->live_host[r >> 5] >> (r & 31)) & 1u) src.push_back(uint32_t(r));
    const uint64_t nl = src.size();
    uint64_t first = 0;
    while (first < nl && src[first] == first) first++;
    cudaStream_t st = ix->mut_stream;
    const size_t row_bytes = size_t(ix->ld) * 4;
    const uint64_t win = std::max<uint64_t>(1, std::min<uint64_t>(nl - first, (size_t(256) << 20) / row_bytes));
    float* scratch = nullptr;
    uint32_t* src_dev = nullptr;
    if (nl > first) {
        CU_OK(cudaMalloc(&scratch, win * row_bytes));
        cudaError_t e = cudaMalloc(&src_dev, win * 4);
        if (e != cudaSuccess) {
            cudaFree(scratch);
            return fail(MVDB_ERR_OOM, "cudaMalloc failed: %s", cudaGetErrorString(e));
        }
    }
    float* mat = static_cast<float*>(ix->mat.ptr());
    int rc = MVDB_OK;
    for (uint64_t a = first; a < nl && rc == MVDB_OK; a += win) {
        uint64_t m = std::min(win, nl - a);
        // destination rows [a, a+m) never overlap the sources of later windows
        // (src[i] >= i), and the gather of this window completes before its copy.
        cudaError_t e = cudaMemcpyAsync(src_dev, src.data() + a, m * 4, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            unsigned grid = unsigned(std::min<uint64_t>((m * 32 + 255) / 256, uint64_t(ix->sm_count) * 16));
            gather_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(mat), reinterpret_cast<float4*>(scratch),
                                                     src_dev, m, ix->ld4);
            LAUNCHED();
            e = cudaMemcpyAsync(mat + a * ix->ld, scratch, m * row_bytes, cudaMemcpyDeviceToDevice, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // src.data() window reuse + error check
        if (e != cudaSuccess) rc = fail(MVDB_ERR_CUDA, "compaction failed: %s", cudaGetErrorString(e));
    }
    cudaFree(scratch);
    cudaFree(src_dev);
    if (rc != MVDB_OK) return rc;
    // rebuild the bitmask: rows [0, nl) live, the rest clear
    size_t old_words = (nt + 31) / 32, new_words = (nl + 31) / 32;
    ix->live_host.assign(old_words, 0u);
    for (uint64_t w = 0; w < nl / 32; w++) ix->live_host[w] = 0xFFFFFFFFu;
    if (nl & 31) ix->live_host[nl / 32] = (1u << (nl & 31)) - 1u;
    if (old_words) CU_OK(cudaMemcpyAsync(ix->live.ptr(), ix->live_host.data(), old_words * 4, cudaMemcpyHostToDevice, st));
    CU_OK(cudaStreamSynchronize(st));
    ix->live_host.resize(new_words);
    ix->ntotal.store(nl, std::memory_order_release);
    ix->ndead.store(0, std::memory_order_release);
    {
        std::lock_guard<std::mutex> sg(ix->shadow_mu);
        ix->shadow_rows = std::min<uint64_t>(ix->shadow_rows, first);  // rows past the first moved one are stale
    }
    {
        std::lock_guard<std::mutex> sg(ix->shadow8_mu);
        ix->shadow8_rows = std::min<uint64_t>(ix->shadow8_rows, first);
    }
    if (ntotal_out) *ntotal_out = int64_t(nl);
    return MVDB_OK;
}

int mvdb_index_dim(const mvdb_index* ix, int* d) {
    if (!ix || !d) return fail(MVDB_ERR_ARG, "null argument");
    *d = ix->d;
    return MVDB_OK;
}

int mvdb_index_ntotal(const mvdb_index* ix, int64_t* ntotal, int64_t* nlive) {
    if (!ix) return fail(MVDB_ERR_ARG, "null index");
    uint64_t nt = ix->ntotal.load(), nd = ix->ndead.load();
    if (ntotal) *ntotal = int64_t(nt);
    if (nlive) *nlive = int64_t(nt - std::min(nd, nt));
    return MVDB_OK;
}

int mvdb_index_reconstruct_n(mvdb_index* ix, int64_t row0, uint64_t n, float* out) {
    ENTER(ix);
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    std::shared_lock<std::shared_mutex> mv(ix->move_mu);
    const uint64_t nt = ix->ntotal.load(std::memory_order_acquire);
    if (row0 < 0 || uint64_t(row0) + n > nt) return fail(MVDB_ERR_ARG, "rows [%lld, +%llu) out of range", (long long)row0, (unsigned long long)n);
    if (n == 0) return MVDB_OK;
    const float* src = static_cast<const float*>(ix->mat.ptr()) + uint64_t(row0) * ix->ld;
    CU_OK(cudaMemcpy2D(out, size_t(ix->d) * 4, src, size_t(ix->ld) * 4, size_t(ix->d) * 4, n, cudaMemcpyDeviceToHost));
    return MVDB_OK;
}
int mvdb_index_reconstruct(mvdb_index* ix, int64_t row, float* out) { return mvdb_index_reconstruct_n(ix, row, 1, out); }

int mvdb_index_device_view(mvdb_index* ix, const float** matrix_dev, int64_t* ld, const uint32_t** live_dev) {
    if (!ix) return fail(MVDB_ERR_ARG, "null index");
    if (matrix_dev) *matrix_dev = static_cast<const float*>(ix->mat.ptr());
    if (ld) *ld = ix->ld;
    if (live_dev) *live_dev = static_cast<const uint32_t*>(ix->live.ptr());
    return MVDB_OK;
}

int mvdb_index_workspace_create(mvdb_index* ix, mvdb_workspace** out) {
    ENTER(ix);
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    RC_OK(ws_new(ix, out));
    std::lock_guard<std::mutex> g(ix->child_mu);
    ix->child_workspaces.insert(*out);
    return MVDB_OK;
}
int mvdb_index_workspace_destroy(mvdb_workspace* ws) {
    if (!ws) return MVDB_OK;
    if (!ws->ix) {   // orphan: the index went first and took the resources with it
        delete ws;
        return MVDB_OK;
    }
    DeviceGuard guard(ws->ix->device);
    {
        std::lock_guard<std::mutex> g(ws->ix->child_mu);
        ws->ix->child_workspaces.erase(ws);
    }
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    ws_free(ws);
    return MVDB_OK;
}

int mvdb_index_search_device(mvdb_index* ix, mvdb_workspace* ws, const float* q_dev, int64_t nq, int64_t k,
                             const uint32_t* mask_dev, uint64_t mask_rows, int normalize_queries,
                             int64_t label_offset, float* D_dev, int64_t* I_dev, void* stream) {
    ENTER(ix);
    if (!ws || ws->ix != ix) return fail(MVDB_ERR_ARG, "workspace does not belong to this index");
    if (nq < 0 || k <= 0) return fail(MVDB_ERR_ARG, "need nq >= 0 and k > 0 (got nq=%lld k=%lld)", (long long)nq, (long long)k);
    if (nq && (!q_dev || !D_dev || !I_dev)) return fail(MVDB_ERR_ARG, "null buffer");
    std::shared_lock<std::shared_mutex> mv(ix->move_mu);
    tl_allow_pdl = true;
    const int rc = run_search(ix, ws, q_dev, nq, k, mask_dev, mask_rows, normalize_queries, label_offset, D_dev, I_dev,
                              static_cast<cudaStream_t>(stream), nullptr);
    tl_allow_pdl = false;
    return rc;
}

// record "this mask's latest writer has been enqueued" on the filter stream
static int mask_mark_ready(mvdb_mask* m) {
    if (!m->ready) CU_OK(cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming));
    CU_OK(cudaEventRecord(m->ready, m->ix->filter_stream));
    return MVDB_OK;
}
// make `st` wait for the kernels that produced the mask (no host synchronisation)
static int mask_wait(const mvdb_mask* m, cudaStream_t st) {
    if (m && m->ready) CU_OK(cudaStreamWaitEvent(st, m->ready, 0));
    return MVDB_OK;
}

// One host-buffer search on its own workspace: stage query (+mask), run, copy results back.
// Host buffers in, host buffers out.  Single queries on the scan kernels (the latency path) avoid both copy-engine
// transfers: the inputs are pulled from pinned memory by a small grid that the scan depends on programmatically
// (pull_stage_kernel), and the kernels write the k results straight into pinned host memory.  Everything else
// (query batches, large k) stages by cudaMemcpyAsync as before.  Option "host_path" switches the pieces off (A/B).
static int search_host_direct(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const uint8_t* mask,
                              uint64_t mask_rows, int normalize_queries, float* D, int64_t* I,
                              const mvdb_mask* handle = nullptr) {
    mvdb_workspace* ws = nullptr;
    RC_OK(pool_acquire(ix, &ws));
    struct Release {
        mvdb_index* ix;
        mvdb_workspace* ws;
        ~Release() { pool_release(ix, ws); }
    } rel{ix, ws};
    std::shared_lock<std::shared_mutex> mv(ix->move_mu);
    cudaStream_t st = ws->stream;
    const size_t qn = size_t(nq) * ix->d, on = size_t(nq) * k;
    const bool single = nq == 1 && k <= ix->fused_k_max;   // answered by one scan kernel that writes its k results once
    const bool zc_out = single && (ix->host_path & 1);
    const bool pull = single && (ix->host_path & 2);
    // labels and distances share one buffer ([I | D]) so that the results come back in ONE transfer
    const size_t out_elems = on + (on + 1) / 2;
    RC_OK(grow_pin(&ws->I_pin, &ws->I_pin_cap, out_elems));
    if (!zc_out) RC_OK(grow_dev(&ws->I_dev, &ws->I_cap, out_elems));
    int64_t* const I_out = zc_out ? ws->I_pin : ws->I_dev;   // pinned host memory is device-addressable under UVA
    float* const D_out = reinterpret_cast<float*>(I_out + on);
    const float* const D_pin = reinterpret_cast<const float*>(ws->I_pin + on);

    // staging layout: [filter words | pad to 128 B | q] in ONE buffer, or the query alone
    const bool stage_mask = !handle && mask != nullptr;
    uint64_t rows = 0;
    size_t words = 0, bytes = 0, q_at = 0;
    if (stage_mask) {
        rows = std::min<uint64_t>(mask_rows, ix->ntotal.load(std::memory_order_acquire));
        mask_rows = rows;
        words = (rows + 31) / 32;
        bytes = (rows + 7) / 8;
        q_at = align_up(words * 4, 128) / 4;
    }
    const size_t total = align_up(q_at + qn, 4);   // whole 16-byte vectors
    uint32_t *pin = nullptr, *dev = nullptr;
    if (stage_mask) {
        RC_OK(grow_dev(&ws->mask_dev, &ws->mask_cap, total));
        RC_OK(grow_pin(&ws->mask_pin, &ws->mask_pin_cap, total));
        pin = ws->mask_pin;
        dev = ws->mask_dev;
    } else {
        RC_OK(grow_dev(&ws->q_dev, &ws->q_cap, total));
        RC_OK(grow_pin(&ws->q_pin, &ws->q_pin_cap, total));
        pin = reinterpret_cast<uint32_t*>(ws->q_pin);
        dev = reinterpret_cast<uint32_t*>(ws->q_dev);
    }
    // A filter the caller already keeps in pinned (page-locked, device-addressable) memory is pulled from where it
    // lies: only its last partial 16-byte vector -- whose spare bits must read as zero -- goes through the staging buffer.
    size_t direct_vecs = 0;
    if (pull && stage_mask && bytes >= 4096 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0) {
        cudaPointerAttributes at = {};
        if (cudaPointerGetAttributes(&at, mask) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer == mask)
            direct_vecs = ((rows & 7) ? bytes - 1 : bytes) / 16;   // a last byte with spare bits is cleaned in the staging buffer
        else
            (void)cudaGetLastError();
    }
    auto stage_filter = [&]() {
        if (!words) return;
        pin[words - 1] = 0;
        const size_t skip = direct_vecs * 16;
        memcpy(reinterpret_cast<uint8_t*>(pin) + skip, mask + skip, bytes - skip);
        if (rows & 7) reinterpret_cast<uint8_t*>(pin)[bytes - 1] &= uint8_t((1u << (rows & 7)) - 1u);
    };
    const uint32_t* mask_dev = nullptr;
    if (handle) {
        mask_dev = handle->dev;      // already resident: no per-query upload
        mask_rows = handle->rows;
        RC_OK(mask_wait(handle, st));
    } else if (stage_mask) {
        mask_dev = dev;
    }
    const float* q_dev = reinterpret_cast<const float*>(dev + q_at);
    memcpy(pin + q_at, q, qn * 4);

    stage_filter();
    if (pull) {
        const uint32_t nvec = uint32_t(total / 4);
        const unsigned grid = unsigned(std::min<uint32_t>((nvec + 255) / 256, uint32_t(ix->sm_count) * 2));
        pull_stage_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(pin), reinterpret_cast<const uint4*>(mask),
                                              uint32_t(direct_vecs), reinterpret_cast<uint4*>(dev), nvec);
        LAUNCHED();
        tl_stage_dep = true;
    } else {
        CU_OK(cudaMemcpyAsync(dev, pin, total * 4, cudaMemcpyHostToDevice, st));
    }
    tl_host_checks_i8 = true;
    int src = run_search(ix, ws, q_dev, nq, k, mask_dev, mask_rows, normalize_queries, 0, D_out, I_out, st, nullptr);
    tl_host_checks_i8 = false;
    tl_stage_dep = false;
    RC_OK(src);
    if (!zc_out) CU_OK(cudaMemcpyAsync(ws->I_pin, ws->I_dev, on * 12, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaStreamSynchronize(st));
    if (ws->i8_ovf_pin && *reinterpret_cast<volatile unsigned int*>(ws->i8_ovf_pin)) {
        // int8 shadow / survivor mode: the list overflowed -- this query is answered by the classic fp32 scan
        // (its inputs are already staged in device memory)
        *ws->i8_ovf_pin = 0u;
        tl_force_scan = true;
        src = run_search(ix, ws, q_dev, nq, k, mask_dev, mask_rows, normalize_queries, 0, D_out, I_out, st, nullptr);
        tl_force_scan = false;
        RC_OK(src);
        if (!zc_out) CU_OK(cudaMemcpyAsync(ws->I_pin, ws->I_dev, on * 12, cudaMemcpyDeviceToHost, st));
        CU_OK(cudaStreamSynchronize(st));
    }
    memcpy(D, D_pin, on * 4);
    memcpy(I, ws->I_pin, on * 8);
    return MVDB_OK;
}

// B concurrent single-query requests (same k, same normalise flag, each with its own optional
// filter) as ONE pass over the matrix: <= 8 queries per scan launch with per-query masks, or --
// when none is filtered and there are >= batch_min_nq of them -- one tensor-core batch.
static int exec_coalesced(mvdb_index* ix, std::vector<CoalesceReq*>& batch) {
    const int64_t B = int64_t(batch.size());
    if (B == 1) {
        CoalesceReq* r = batch[0];
        return search_host_direct(ix, r->q, 1, r->k, r->mask, r->mask_rows, r->normalize, r->D, r->I, r->handle);
    }
    const int64_t k = batch[0]->k;
    mvdb_workspace* ws = nullptr;
    RC_OK(pool_acquire(ix, &ws));
    struct Release {
        mvdb_index* ix;
        mvdb_workspace* ws;
        ~Release() { pool_release(ix, ws); }
    } rel{ix, ws};
    std::shared_lock<std::shared_mutex> mv(ix->move_mu);
    cudaStream_t st = ws->stream;
    const uint64_t n = ix->ntotal.load(std::memory_order_acquire);
    const size_t qn = size_t(B) * ix->d, on = size_t(B) * k, words = (n + 31) / 32;
    int n_masked = 0;   // requests whose filter still has to be staged from host memory
    bool any_filter = false;
    for (auto* r : batch) {
        n_masked += (r->mask != nullptr && !r->handle);
        any_filter |= r->mask != nullptr || r->handle != nullptr;
    }
    RC_OK(grow_dev(&ws->q_dev, &ws->q_cap, qn));
    RC_OK(grow_pin(&ws->q_pin, &ws->q_pin_cap, qn));
    const size_t out_elems = on + (on + 1) / 2;   // [I | D] in one buffer: one D2H transfer
    RC_OK(grow_dev(&ws->I_dev, &ws->I_cap, out_elems));
    RC_OK(grow_pin(&ws->I_pin, &ws->I_pin_cap, out_elems));
    float* const D_dev = reinterpret_cast<float*>(ws->I_dev + on);
    const float* const D_pin = reinterpret_cast<const float*>(ws->I_pin + on);
    if (n_masked) {
        RC_OK(grow_dev(&ws->mask_dev, &ws->mask_cap, std::max<size_t>(words * n_masked, 1)));
        RC_OK(grow_pin(&ws->mask_pin, &ws->mask_pin_cap, std::max<size_t>(words * n_masked, 1)));
    }
    std::vector<QMaskRef> qmasks(size_t(B), QMaskRef{nullptr, 0u});
    int slot = 0;
    for (int64_t i = 0; i < B; i++) {
        CoalesceReq* r = batch[size_t(i)];
        memcpy(ws->q_pin + i * ix->d, r->q, size_t(ix->d) * 4);
        if (r->handle) {
            qmasks[size_t(i)] = QMaskRef{r->handle->dev, r->handle->words};
            RC_OK(mask_wait(r->handle, st));
        } else if (r->mask) {
            // rows past the caller's mask_rows are not admissible: zero-filled tail
            uint32_t* dst = ws->mask_pin + size_t(slot) * words;
            const uint64_t rows = std::min<uint64_t>(r->mask_rows, n);
            const size_t bytes = (rows + 7) / 8;
            memset(dst, 0, words * 4);
            memcpy(dst, r->mask, bytes);
            if (rows & 7) reinterpret_cast<uint8_t*>(dst)[bytes - 1] &= uint8_t((1u << (rows & 7)) - 1u);
            qmasks[size_t(i)] = QMaskRef{ws->mask_dev + size_t(slot) * words, uint32_t(words)};
            slot++;
        }
    }
    CU_OK(cudaMemcpyAsync(ws->q_dev, ws->q_pin, qn * 4, cudaMemcpyHostToDevice, st));
    if (n_masked && words)
        CU_OK(cudaMemcpyAsync(ws->mask_dev, ws->mask_pin, words * 4 * n_masked, cudaMemcpyHostToDevice, st));
    RC_OK(run_search(ix, ws, ws->q_dev, B, k, nullptr, 0, batch[0]->normalize, 0, D_dev, ws->I_dev, st, nullptr,
                     any_filter ? qmasks.data() : nullptr));
    CU_OK(cudaMemcpyAsync(ws->I_pin, ws->I_dev, on * 12, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < B; i++) {
        memcpy(batch[size_t(i)]->D, D_pin + i * k, size_t(k) * 4);
        memcpy(batch[size_t(i)]->I, ws->I_pin + i * k, size_t(k) * 8);
    }
    return MVDB_OK;
}

// Leader/follower coalescing with zero added latency when idle: a thread that finds a free leader
// seat (one or two, see below) takes it and serves whatever is queued (including its own request);
// while a batch runs on the GPU new arrivals queue up and form the next batch.
static int coalesced_search(mvdb_index* ix, CoalesceReq& req) {
    std::unique_lock<std::mutex> lk(ix->co_mu);
    ix->co_queue.push_back(&req);
    bool seat;   // does this thread hold one of the leader seats?
    // Two leaders overlap one batch's host staging with the other's GPU pass -- worth it only while a
    // pass is as short as the staging.  Once the matrix pass dominates (>= 256 MB: tens of us), two
    // concurrent passes just share the HBM bandwidth and halve the batch size; one leader then lets
    // the queue grow into one bigger batch per pass.
    int seats = ix->co_max_leaders;
    if (seats <= 0)
        seats = (ix->ntotal.load(std::memory_order_acquire) * uint64_t(ix->ld) * 4ull >= (256ull << 20)) ? 1 : 2;
    if (ix->co_leaders >= seats) {
        req.cv.wait(lk, [&] { return req.done || req.promote; });
        seat = req.promote;   // a departing leader handed its seat over (possibly after our request was served)
    } else {
        ix->co_leaders++;
        seat = true;
    }
    if (seat) {
        while (!req.done) {
            // Callers that were served by one shared pass come back TOGETHER (within the time their host code
            // needs to issue the next call): the first of them to arrive would otherwise run a pass for itself
            // alone while the others queue up behind it, and every round of B callers would cost two passes.  So a
            // leader that can expect company (the previous batch had several queries) gives it a moment --
            // nothing measurable for small indexes, a fraction of a millisecond when a pass takes milliseconds.
            if (ix->co_wait_pct > 0 && ix->co_last_batch + ix->co_prev_batch > 2 && !ix->co_queue.empty() &&
                ix->co_queue.size() < ix->co_last_batch + ix->co_prev_batch) {
                // wait while callers keep arriving: until the previous company is back, or nobody has arrived for
                // a short gap (2 % of a pass, >= 20 us), or co_wait_pct % of a pass has gone by
                const double pass_ns = double(ix->ntotal.load(std::memory_order_acquire)) * double(ix->ld) * 2.0 / 6.0e12 * 1e9;
                const auto t_start = std::chrono::steady_clock::now();
                const auto deadline = t_start + std::chrono::nanoseconds(int64_t(pass_ns * ix->co_wait_pct / 100.0));
                const auto gap = std::chrono::nanoseconds(std::max<int64_t>(20000, int64_t(pass_ns * 0.02)));
                const size_t want = std::min(ix->co_last_batch + ix->co_prev_batch, size_t(ix->coalesce_max));
                size_t seen = ix->co_queue.size();
                auto last_growth = t_start;
                for (;;) {
                    const auto now = std::chrono::steady_clock::now();
                    if (ix->co_queue.size() > seen) {
                        seen = ix->co_queue.size();
                        last_growth = now;
                    }
                    if (seen >= want || now >= deadline || now - last_growth > gap) break;
                    lk.unlock();
                    std::this_thread::yield();
                    lk.lock();
                }
            }
            if (ix->co_queue.empty()) {
                // our own request is in flight inside the other leader's batch
                req.cv.wait(lk, [&] { return req.done; });
                break;
            }
            // batch = the queue head plus every compatible request behind it (same k, same
            // normalise flag): up to coalesce_max unfiltered ones, or up to 8 once a filter is in
            std::vector<CoalesceReq*> batch;
            CoalesceReq* head = ix->co_queue.front();
            const bool single = head->k > ix->fused_k_max;   // large-k path serves one query at a time
            bool masked = false;
            for (auto it = ix->co_queue.begin(); it != ix->co_queue.end();) {
                CoalesceReq* r = *it;
                // filters that must be staged from host memory limit the batch to one scan launch;
                // unfiltered queries and device-resident filters (mask handles) batch freely
                const bool would_mask = masked || (r->mask != nullptr && r->handle == nullptr);
                const size_t cap = single ? 1 : (would_mask ? 8 : size_t(ix->coalesce_max));
                if (r->k == head->k && r->normalize == head->normalize && batch.size() < cap) {
                    batch.push_back(r);
                    masked = would_mask;
                    it = ix->co_queue.erase(it);
                } else {
                    ++it;
                }
            }
            ix->co_prev_batch = ix->co_last_batch;
            ix->co_last_batch = batch.size();
            lk.unlock();
            int rc = exec_coalesced(ix, batch);
            std::string err = rc != MVDB_OK ? g_err : std::string();
            lk.lock();
            for (CoalesceReq* r : batch) {
                r->rc = rc;
                r->err = err;
                r->done = true;
                if (r != &req) r->cv.notify_one();
            }
        }
        // leave the seat to the first waiter that has not been promoted yet, else free it
        CoalesceReq* next = nullptr;
        for (CoalesceReq* r : ix->co_queue)
            if (!r->promote) {
                next = r;
                break;
            }
        if (next) {
            next->promote = true;
            next->cv.notify_one();
        } else {
            ix->co_leaders--;
        }
    }
    if (req.rc != MVDB_OK) g_err = req.err;
    return req.rc;
}

int mvdb_index_mask_create(mvdb_index* ix, const uint8_t* mask, uint64_t mask_rows, mvdb_mask** out) {
    ENTER(ix);
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    *out = nullptr;
    if (!mask && mask_rows) return fail(MVDB_ERR_ARG, "null mask");
    if (mask_rows > 0xFFFFFFF0ull) return fail(MVDB_ERR_ARG, "mask too long");
    const size_t words = (mask_rows + 31) / 32, bytes = (mask_rows + 7) / 8;
    std::vector<uint32_t> host(std::max<size_t>(words, 1), 0u);   // zero tail: rows past mask_rows are not admissible
    if (bytes) memcpy(host.data(), mask, bytes);
    if (mask_rows & 7) reinterpret_cast<uint8_t*>(host.data())[bytes - 1] &= uint8_t((1u << (mask_rows & 7)) - 1u);
    mvdb_mask* m = new mvdb_mask{ix, nullptr, mask_rows, uint32_t(words)};
    cudaError_t e = cudaMalloc(&m->dev, host.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice, ix->filter_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->filter_stream);   // `host` goes out of scope; the upload is complete
    if (e != cudaSuccess) {
        cudaFree(m->dev);
        delete m;
        return fail(e == cudaErrorMemoryAllocation ? MVDB_ERR_OOM : MVDB_ERR_CUDA, "mask upload failed: %s", cudaGetErrorString(e));
    }
    {
        std::lock_guard<std::mutex> g(ix->child_mu);
        ix->child_masks.insert(m);
    }
    *out = m;
    return MVDB_OK;
}

int mvdb_mask_destroy(mvdb_mask* m) {
    if (!m) return MVDB_OK;
    if (!m->ix) {   // orphan: the index was destroyed first and released the device memory
        delete m;
        return MVDB_OK;
    }
    {
        std::lock_guard<std::mutex> g(m->ix->child_mu);
        m->ix->child_masks.erase(m);
    }
    DeviceGuard guard(m->ix->device);
    cudaFree(m->dev);   // synchronises the device: no filter kernel or search still reads it
    if (m->ready) cudaEventDestroy(m->ready);
    delete m;
    return MVDB_OK;
}

static int search_entry(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const uint8_t* mask, uint64_t mask_rows,
                        const mvdb_mask* handle, int normalize_queries, float* D, int64_t* I);

// ---- device-side filter evaluation -------------------------------------------------------
int mvdb_column_create(mvdb_index* ix, mvdb_column** out) {
    if (!ix || !out) return fail(MVDB_ERR_ARG, "null argument");
    *out = new mvdb_column();
    (*out)->ix = ix;
    std::lock_guard<std::mutex> g(ix->child_mu);
    ix->child_columns.insert(*out);
    return MVDB_OK;
}

int mvdb_column_destroy(mvdb_column* c) {
    if (!c) return MVDB_OK;
    if (!c->ix) {   // orphan
        delete c;
        return MVDB_OK;
    }
    {
        std::lock_guard<std::mutex> g(c->ix->child_mu);
        c->ix->child_columns.erase(c);
    }
    DeviceGuard guard(c->ix->device);
    cudaFree(c->vals);
    cudaFree(c->has);
    delete c;
    return MVDB_OK;
}

int mvdb_column_append(mvdb_column* c, const double* values, const uint8_t* present, uint64_t n) {
    if (!c) return fail(MVDB_ERR_ARG, "null column");
    if (!c->ix) return fail(MVDB_ERR_STATE, "the column's index has been destroyed");
    if (n == 0) return MVDB_OK;
    if (!values || !present) return fail(MVDB_ERR_ARG, "null data");
    ENTER(c->ix);
    // everything on the filter stream, so that it is ordered with the predicate kernels that read the
    // column (searches run on non-blocking streams: the legacy default stream orders nothing for them)
    cudaStream_t fs = c->ix->filter_stream;
    const uint64_t need = c->len + n;
    if (need > c->cap) {
        const uint64_t ncap = align_up(std::max<uint64_t>(need, c->cap * 2), 1024);
        double* nv = nullptr;
        uint32_t* nh = nullptr;
        CU_OK(cudaMalloc(&nv, ncap * 8));
        cudaError_t e = cudaMalloc(&nh, ncap / 8 + 8);
        if (e == cudaSuccess) e = cudaMemsetAsync(nh, 0, ncap / 8 + 8, fs);
        if (e == cudaSuccess && c->len) e = cudaMemcpyAsync(nv, c->vals, c->len * 8, cudaMemcpyDeviceToDevice, fs);
        if (e == cudaSuccess && c->len) e = cudaMemcpyAsync(nh, c->has, (c->len + 31) / 32 * 4, cudaMemcpyDeviceToDevice, fs);
        if (e == cudaSuccess) e = cudaStreamSynchronize(fs);   // the old buffers are freed below
        if (e != cudaSuccess) {
            cudaFree(nv);
            cudaFree(nh);
            return fail(MVDB_ERR_CUDA, "column growth failed: %s", cudaGetErrorString(e));
        }
        cudaFree(c->vals);
        cudaFree(c->has);
        c->vals = nv;
        c->has = nh;
        c->cap = ncap;
    }
    CU_OK(cudaMemcpyAsync(c->vals + c->len, values, n * 8, cudaMemcpyHostToDevice, fs));
    // presence bits: merge the partially filled first word on the host
    const uint64_t w0 = c->len / 32, w1 = (need + 31) / 32;
    std::vector<uint32_t> words(size_t(w1 - w0), 0u);
    if (c->len % 32) {
        CU_OK(cudaMemcpyAsync(words.data(), c->has + w0, 4, cudaMemcpyDeviceToHost, fs));
        CU_OK(cudaStreamSynchronize(fs));
    }
    for (uint64_t i = 0; i < n; i++)
        if (present[i]) {
            const uint64_t r = c->len + i;
            words[size_t(r / 32 - w0)] |= 1u << (r % 32);
        }
    CU_OK(cudaMemcpyAsync(c->has + w0, words.data(), words.size() * 4, cudaMemcpyHostToDevice, fs));
    CU_OK(cudaStreamSynchronize(fs));   // the caller's and our host buffers are free again
    c->len = need;
    return MVDB_OK;
}

static int new_mask(mvdb_index* ix, uint64_t rows, mvdb_mask** out) {
    const size_t words = std::max<size_t>((rows + 31) / 32, 1);
    mvdb_mask* m = new mvdb_mask{ix, nullptr, rows, uint32_t((rows + 31) / 32)};
    cudaError_t e = cudaMalloc(&m->dev, words * 4);
    if (e != cudaSuccess) {
        delete m;
        return fail(MVDB_ERR_OOM, "mask allocation failed: %s", cudaGetErrorString(e));
    }
    {
        std::lock_guard<std::mutex> g(ix->child_mu);
        ix->child_masks.insert(m);
    }
    *out = m;
    return MVDB_OK;
}

int mvdb_mask_from_predicate(mvdb_index* ix, const mvdb_column* c, int op, double operand, mvdb_mask** out) {
    ENTER(ix);
    if (!c || !out || c->ix != ix) return fail(MVDB_ERR_ARG, "bad column");
    if (op < 0 || op > 5) return fail(MVDB_ERR_ARG, "bad operator");
    *out = nullptr;
    mvdb_mask* m = nullptr;
    RC_OK(new_mask(ix, c->len, &m));
    if (m->words) {
        predicate_mask_kernel<<<(m->words + 255) / 256, 256, 0, ix->filter_stream>>>(c->vals, c->has, c->len, op, operand, m->dev,
                                                                                      m->words);
        LAUNCHED();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || mask_mark_ready(m) != MVDB_OK) {
        mvdb_mask_destroy(m);
        return fail(MVDB_ERR_CUDA, "predicate kernel failed: %s", cudaGetErrorString(e));
    }
    *out = m;
    return MVDB_OK;
}

int mvdb_mask_create_filled(mvdb_index* ix, uint64_t rows, mvdb_mask** out) {
    ENTER(ix);
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    *out = nullptr;
    mvdb_mask* m = nullptr;
    RC_OK(new_mask(ix, rows, &m));
    if (m->words) {
        fill_mask_kernel<<<(m->words + 255) / 256, 256, 0, ix->filter_stream>>>(m->dev, rows, m->words);
        LAUNCHED();
    }
    if (cudaGetLastError() != cudaSuccess || mask_mark_ready(m) != MVDB_OK) {
        mvdb_mask_destroy(m);
        return fail(MVDB_ERR_CUDA, "fill kernel failed");
    }
    *out = m;
    return MVDB_OK;
}

int mvdb_mask_combine(mvdb_mask* dst, const mvdb_mask* src, int how) {
    if (!dst || !src || dst->ix != src->ix || !dst->ix) return fail(MVDB_ERR_ARG, "bad masks");
    if (how < 0 || how > 2) return fail(MVDB_ERR_ARG, "bad combine mode");
    ENTER(dst->ix);
    if (dst->words) {
        // src was produced on the same stream (or uploaded synchronously): stream order is enough
        combine_mask_kernel<<<(dst->words + 255) / 256, 256, 0, dst->ix->filter_stream>>>(dst->dev, dst->words, src->dev, src->words,
                                                                                           how);
        LAUNCHED();
    }
    CU_OK(cudaGetLastError());
    return mask_mark_ready(dst);
}

int mvdb_mask_count(const mvdb_mask* m, uint64_t* count) {
    if (!m || !count) return fail(MVDB_ERR_ARG, "null argument");
    if (!m->ix) return fail(MVDB_ERR_STATE, "the mask's index has been destroyed");
    mvdb_index* ix = m->ix;
    ENTER(ix);
    // one 8-byte counter per index, serialised by a mutex (a cudaMalloc per count would dominate)
    std::lock_guard<std::mutex> cg(ix->count_mu);
    if (!ix->count_dev) CU_OK(cudaMalloc(&ix->count_dev, 8));
    unsigned long long* dev = ix->count_dev;
    cudaStream_t fs = ix->filter_stream;   // behind the kernels that built the mask
    cudaError_t e = cudaMemsetAsync(dev, 0, 8, fs);
    // rows of the mask that are still live; the live bitmask covers at least as many words
    const uint64_t nt = ix->ntotal.load(std::memory_order_acquire);
    const uint32_t words = uint32_t(std::min<uint64_t>(m->words, (nt + 31) / 32));
    if (e == cudaSuccess && words) {
        const uint32_t* live = ix->ndead.load(std::memory_order_acquire) ? static_cast<const uint32_t*>(ix->live.ptr()) : nullptr;
        count_mask_kernel<<<(words + 255) / 256, 256, 0, fs>>>(m->dev, live, words, dev);
        LAUNCHED();
        e = cudaGetLastError();
    }
    unsigned long long v = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&v, dev, 8, cudaMemcpyDeviceToHost, fs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(fs);
    if (e != cudaSuccess) return fail(MVDB_ERR_CUDA, "mask count failed: %s", cudaGetErrorString(e));
    *count = v;
    return MVDB_OK;
}

int mvdb_index_search_with_mask(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const mvdb_mask* m,
                                int normalize_queries, float* D, int64_t* I) {
    if (m && m->ix != ix) return fail(MVDB_ERR_ARG, "mask handle belongs to another index (or its index has been destroyed)");
    return search_entry(ix, q, nq, k, nullptr, 0, m, normalize_queries, D, I);
}

int mvdb_index_search(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const uint8_t* mask, uint64_t mask_rows,
                      int normalize_queries, float* D, int64_t* I) {
    return search_entry(ix, q, nq, k, mask, mask_rows, nullptr, normalize_queries, D, I);
}

static int search_entry(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const uint8_t* mask, uint64_t mask_rows,
                        const mvdb_mask* handle, int normalize_queries, float* D, int64_t* I) {
    ENTER(ix);
    if (nq < 0 || k <= 0) return fail(MVDB_ERR_ARG, "need nq >= 0 and k > 0 (got nq=%lld k=%lld)", (long long)nq, (long long)k);
    if (nq == 0) return MVDB_OK;
    if (!q || !D || !I) return fail(MVDB_ERR_ARG, "null buffer");
    if (nq == 1 && ix->coalesce) {
        CoalesceReq req;
        req.q = q;
        req.k = k;
        req.mask = mask;
        req.mask_rows = mask_rows;
        req.handle = handle;
        req.normalize = normalize_queries ? 1 : 0;
        req.D = D;
        req.I = I;
        return coalesced_search(ix, req);
    }
    return search_host_direct(ix, q, nq, k, mask, mask_rows, normalize_queries, D, I, handle);
}

int mvdb_exchange_create(int device, int rank, int world, int k_max, int nq_max, mvdb_exchange** out) {
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(MVDB_ERR_ARG, "bad rank/world %d/%d", rank, world);
    if (k_max < 1 || k_max > 128 || nq_max < 1 || nq_max > 8) return fail(MVDB_ERR_ARG, "need 1 <= k_max <= 128, 1 <= nq_max <= 8");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MVDB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    mvdb_exchange* x = new mvdb_exchange();
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->k_max = k_max;
    x->nq_max = nq_max;
    x->recv_words = size_t(2) * world * nq_max * k_max;
    const size_t words = x->recv_words + size_t(2) * world;
    cudaError_t e = cudaMalloc(&x->local, words * 8);
    if (e == cudaSuccess) e = cudaMemset(x->local, 0, words * 8);
    if (e == cudaSuccess) e = cudaMalloc(&x->dev, sizeof(XchgDev));
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&x->status), sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *x->status = 0u;
        e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&x->status_dev), x->status, 0);
    }
    if (e != cudaSuccess) {
        cudaFree(x->local);
        cudaFree(x->dev);
        if (x->status) cudaFreeHost(x->status);
        delete x;
        return fail(MVDB_ERR_CUDA, "exchange allocation failed: %s", cudaGetErrorString(e));
    }
    *out = x;
    return MVDB_OK;
}

int mvdb_exchange_ipc_handle(mvdb_exchange* x, void* handle64) {
    if (!x || !handle64) return fail(MVDB_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(x->device);
    cudaIpcMemHandle_t h;
    CU_OK(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle64, &h, 64);
    return MVDB_OK;
}

static int exchange_upload(mvdb_exchange* x) {
    CU_OK(cudaMemcpy(x->dev, &x->host, sizeof(XchgDev), cudaMemcpyHostToDevice));
    return MVDB_OK;
}

int mvdb_exchange_connect(mvdb_exchange* x, const void* handles, const int64_t* offsets) {
    if (!x || !handles || !offsets) return fail(MVDB_ERR_ARG, "null argument");
    DeviceGuard guard(x->device);
    if (!guard.ok) return fail(MVDB_ERR_CUDA, "cudaSetDevice(%d) failed", x->device);
    for (int p = 0; p < x->world; p++) {
        void* base = x->local;
        if (p != x->rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const char*>(handles) + size_t(p) * 64, 64);
            CU_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            x->peer_base[p] = base;
        }
        x->host.recv[p] = static_cast<uint64_t*>(base);
        x->host.flags[p] = static_cast<uint64_t*>(base) + x->recv_words;
        x->host.offsets[p] = offsets[p];
    }
    x->host.world = x->world;
    x->host.rank = x->rank;
    x->host.k_max = x->k_max;
    x->host.nq_max = x->nq_max;
    x->host.status = x->status_dev;
    x->host.timeout_ns = x->timeout_ns;
    RC_OK(exchange_upload(x));
    x->connected = true;
    return MVDB_OK;
}

/* Same-process flavour: the `n` exchange objects (ranks 0..n-1, one per device) live in THIS process, so
 * the peers' buffers are reached through cudaDeviceEnablePeerAccess instead of CUDA IPC handles. */
int mvdb_exchange_connect_local(mvdb_exchange* const* xs, int n, const int64_t* offsets) {
    if (!xs || !offsets || n < 1 || n > kMaxWorld) return fail(MVDB_ERR_ARG, "bad arguments");
    for (int i = 0; i < n; i++)
        if (!xs[i] || xs[i]->world != n || xs[i]->rank != i) return fail(MVDB_ERR_ARG, "exchange %d is not rank %d of %d", i, i, n);
    for (int i = 0; i < n; i++) {
        mvdb_exchange* x = xs[i];
        DeviceGuard guard(x->device);
        if (!guard.ok) return fail(MVDB_ERR_CUDA, "cudaSetDevice(%d) failed", x->device);
        for (int p = 0; p < n; p++) {
            if (p != i && xs[p]->device != x->device) {
                int can = 0;
                CU_OK(cudaDeviceCanAccessPeer(&can, x->device, xs[p]->device));
                if (!can) return fail(MVDB_ERR_STATE, "device %d cannot access device %d", x->device, xs[p]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(xs[p]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
                else if (e != cudaSuccess) return fail(MVDB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", xs[p]->device, cudaGetErrorString(e));
            }
            x->host.recv[p] = xs[p]->local;
            x->host.flags[p] = xs[p]->local + xs[p]->recv_words;
            x->host.offsets[p] = offsets[p];
        }
        x->host.world = n;
        x->host.rank = i;
        x->host.k_max = x->k_max;
        x->host.nq_max = x->nq_max;
        x->host.status = x->status_dev;
        x->host.timeout_ns = x->timeout_ns;
        RC_OK(exchange_upload(x));
        x->connected = true;
        x->local_peers = true;
    }
    return MVDB_OK;
}

int mvdb_exchange_set_option(mvdb_exchange* x, const char* name, int64_t value) {
    if (!x || !name) return fail(MVDB_ERR_ARG, "null argument");
    if (std::string(name) == "timeout_ms") {
        if (value < 1 || value > 600000) return fail(MVDB_ERR_ARG, "timeout_ms must be 1..600000");
        x->timeout_ns = uint64_t(value) * 1000000ull;
        x->host.timeout_ns = x->timeout_ns;
        if (x->connected) {
            DeviceGuard guard(x->device);
            CU_OK(cudaDeviceSynchronize());
            return exchange_upload(x);
        }
        return MVDB_OK;
    }
    return fail(MVDB_ERR_ARG, "unknown exchange option '%s'", name);
}

int mvdb_exchange_set_offsets(mvdb_exchange* x, const int64_t* offsets) {
    if (!x || !offsets) return fail(MVDB_ERR_ARG, "null argument");
    DeviceGuard guard(x->device);
    CU_OK(cudaDeviceSynchronize());
    for (int p = 0; p < x->world; p++) x->host.offsets[p] = offsets[p];
    return exchange_upload(x);
}

int mvdb_exchange_status(mvdb_exchange* x, int* timed_out) {
    if (!x || !timed_out) return fail(MVDB_ERR_ARG, "null argument");
    // the word lives in pinned host memory: valid after any synchronise that covered the search, no CUDA call
    *timed_out = int(*reinterpret_cast<volatile unsigned int*>(x->status));
    return MVDB_OK;
}

int mvdb_exchange_destroy(mvdb_exchange* x) {
    if (!x) return MVDB_OK;
    DeviceGuard guard(x->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < x->world; p++)
        if (x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
    cudaFree(x->local);
    cudaFree(x->dev);
    if (x->status) cudaFreeHost(x->status);
    delete x;
    return MVDB_OK;
}

int mvdb_index_search_exchange(mvdb_index* ix, mvdb_workspace* ws, mvdb_exchange* x, const float* q_dev, int64_t nq,
                               int64_t k, const uint32_t* mask_dev, uint64_t mask_rows, int normalize_queries,
                               float* D_dev, int64_t* I_dev, void* stream) {
    ENTER(ix);
    if (!ws || ws->ix != ix) return fail(MVDB_ERR_ARG, "workspace does not belong to this index");
    if (!x || x->device != ix->device) return fail(MVDB_ERR_ARG, "exchange does not belong to this device");
    if (nq < 0 || k <= 0) return fail(MVDB_ERR_ARG, "need nq >= 0 and k > 0 (got nq=%lld k=%lld)", (long long)nq, (long long)k);
    if (nq && (!q_dev || !D_dev || !I_dev)) return fail(MVDB_ERR_ARG, "null buffer");
    std::shared_lock<std::shared_mutex> mv(ix->move_mu);
    tl_allow_pdl = true;
    const int rc = run_search(ix, ws, q_dev, nq, k, mask_dev, mask_rows, normalize_queries, 0, D_dev, I_dev,
                              static_cast<cudaStream_t>(stream), x);
    tl_allow_pdl = false;
    return rc;
}

// Everything a fused scan of `nq` queries would otherwise set up lazily at launch time (scratch
// allocations, per-kernel shared-memory opt-in): cudaMalloc synchronises the whole device, which must
// not happen once a peer's scan -- possibly on the same device -- is already waiting for ours.
static int prepare_fused_scan(mvdb_index* ix, mvdb_workspace* ws, int64_t nq, int64_t k) {
    RC_OK(ws_scratch(ws));
    RC_OK(grow_dev(&ws->partials, &ws->partials_cap, size_t(8) * size_t(ix->sm_count) * 4 * size_t(k)));
    ScanParams p = {};
    p.n = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(ix->ntotal.load(std::memory_order_acquire), 0xFFFFFFF0ull)));
    p.d = ix->d;
    p.ld4 = ix->ld4;
    p.k = int(k);
    for (int g : {8, 4, 2, 1}) {
        if (nq < g && g != 1) continue;
        ScanPlan plan;
        RC_OK(plan_scan(ix, p, g, &plan));   // cudaFuncSetAttribute also forces the (lazily loaded) kernel in
    }
    // same for the stand-in kernel of an empty shard: with lazy module loading the FIRST launch of a kernel
    // may synchronise the device
    cudaFuncAttributes fa;
    CU_OK(cudaFuncGetAttributes(&fa, xchg_empty_kernel));
    if (i8_eligible(ix, nq, k, p.n)) RC_OK(i8_prepare(ix, ws, p.n, ws->stream));
    if (ix->survivor_tail && k > 32 && k <= 128 && (ix->ld4 + 31) / 32 <= 8) RC_OK(sv_prepare(ix, ws, (ix->ld4 + 31) / 32, ws->stream));
    return MVDB_OK;
}

// ---------------------------------------------------------------------------
// Shard group: the row shards of ONE database on several GPUs of one box, driven by ONE process.
// Stands in for the single in-memory index of the reference's ShardedVectorDatabase
// (ref sharded_vector_database.py:79-84, 598-662) when `devices=[...]` spreads it over GPUs.
// One host thread stages the query (+ each shard's filter) to every device, launches every
// device's scan and only then waits: the scans run concurrently, their last CTAs exchange the
// per-shard top-k over NVLink (peer stores, same kernels as the one-process-per-GPU path) and
// every device ends with the merged global list; the host fetches it from shard 0 in one transfer.
// ---------------------------------------------------------------------------
struct mvdb_group {
    int n = 0;
    std::vector<mvdb_index*> shards;
    std::vector<mvdb_workspace*> ws;
    std::vector<mvdb_exchange*> xch;
    std::vector<float*> D_dev;        // per shard: [nq_cap * k_cap] merged distances (every shard gets the same)
    std::vector<int64_t*> I_dev;
    size_t out_cap = 0;
    std::mutex mu;                    // all shards must see the same sequence of searches
};

int mvdb_group_create(mvdb_index* const* shards, int n, mvdb_group** out) {
    if (!out) return fail(MVDB_ERR_ARG, "null out");
    *out = nullptr;
    if (!shards || n < 1 || n > kMaxWorld) return fail(MVDB_ERR_ARG, "need 1..%d shards", kMaxWorld);
    for (int i = 0; i < n; i++) {
        if (!shards[i]) return fail(MVDB_ERR_ARG, "null shard");
        if (shards[i]->group) return fail(MVDB_ERR_STATE, "shard %d already belongs to a group", i);
        if (shards[i]->d != shards[0]->d) return fail(MVDB_ERR_ARG, "shards differ in dimension");
        for (int j = 0; j < i; j++)
            if (shards[j] == shards[i]) return fail(MVDB_ERR_ARG, "shard %d listed twice", i);
        // several shards MAY share a device (tests on a one-GPU box): a scan whose last CTA waits for a peer
        // holds one SM only, so the peer's scan on the same device still runs
    }
    mvdb_group* g = new mvdb_group();
    g->n = n;
    g->shards.assign(shards, shards + n);
    g->ws.assign(size_t(n), nullptr);
    g->xch.assign(size_t(n), nullptr);
    g->D_dev.assign(size_t(n), nullptr);
    g->I_dev.assign(size_t(n), nullptr);
    int rc = MVDB_OK;
    std::vector<int64_t> offs(size_t(n), 0);
    for (int i = 0; i < n && rc == MVDB_OK; i++) {
        offs[size_t(i)] = int64_t(i) << 40;   // label = shard << 40 | row of the shard (stable under inserts)
        DeviceGuard guard(shards[i]->device);
        rc = ws_new(shards[i], &g->ws[size_t(i)]);
        if (rc == MVDB_OK) rc = mvdb_exchange_create(shards[i]->device, i, n, 128, 8, &g->xch[size_t(i)]);
    }
    if (rc == MVDB_OK) rc = mvdb_exchange_connect_local(g->xch.data(), n, offs.data());
    if (rc != MVDB_OK) {
        std::string keep = g_err;
        mvdb_group_destroy(g);
        g_err = keep;
        return rc;
    }
    for (int i = 0; i < n; i++) shards[i]->group = g;   // a member cannot be destroyed before its group
    *out = g;
    return MVDB_OK;
}

int mvdb_group_destroy(mvdb_group* g) {
    if (!g) return MVDB_OK;
    // nobody unmaps while a peer may still write: drain every device first
    for (int i = 0; i < g->n; i++) {
        DeviceGuard guard(g->shards[size_t(i)]->device);
        cudaDeviceSynchronize();
    }
    for (int i = 0; i < g->n; i++) {
        DeviceGuard guard(g->shards[size_t(i)]->device);
        if (g->shards[size_t(i)]->group == g) g->shards[size_t(i)]->group = nullptr;
        if (g->xch[size_t(i)]) mvdb_exchange_destroy(g->xch[size_t(i)]);
        if (g->ws[size_t(i)]) ws_free(g->ws[size_t(i)]);
        cudaFree(g->D_dev[size_t(i)]);
        cudaFree(g->I_dev[size_t(i)]);
    }
    delete g;
    return MVDB_OK;
}

int mvdb_group_set_option(mvdb_group* g, const char* name, int64_t value) {
    if (!g || !name) return fail(MVDB_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(g->mu);
    for (int i = 0; i < g->n; i++) RC_OK(mvdb_exchange_set_option(g->xch[size_t(i)], name, value));
    return MVDB_OK;
}

int mvdb_group_search(mvdb_group* g, const float* q, int64_t nq, int64_t k, const mvdb_mask* const* masks,
                      const uint8_t* const* host_masks, const uint64_t* host_mask_rows, int normalize_queries,
                      float* D, int64_t* I) {
    if (!g) return fail(MVDB_ERR_ARG, "null group");
    if (nq < 0 || k <= 0) return fail(MVDB_ERR_ARG, "need nq >= 0 and k > 0 (got nq=%lld k=%lld)", (long long)nq, (long long)k);
    if (nq == 0) return MVDB_OK;
    if (!q || !D || !I) return fail(MVDB_ERR_ARG, "null buffer");
    if (k > 128) return fail(MVDB_ERR_ARG, "the fused exchange supports k <= 128 (got %lld)", (long long)k);
    const int d = g->shards[0]->d;
    const size_t qn = size_t(nq) * d, on = size_t(nq) * k;
    std::lock_guard<std::mutex> lk(g->mu);
    int cur_dev = -1;
    cudaGetDevice(&cur_dev);
    struct Restore {
        int dev;
        ~Restore() { if (dev >= 0) cudaSetDevice(dev); }
    } restore{cur_dev};
    // shared locks of every shard for the whole search: rows do not move under any of the scans
    std::vector<std::shared_lock<std::shared_mutex>> locks;
    locks.reserve(size_t(g->n));
    for (int s = 0; s < g->n; s++) locks.emplace_back(g->shards[size_t(s)]->move_mu);
    // phase 1: everything that can fail or block (allocations, staging) BEFORE the first launch -- once one
    // device's scan is in flight every other device must launch too, or the first one waits for its timeout
    std::vector<const uint32_t*> mask_dev(size_t(g->n), nullptr);
    std::vector<uint64_t> mask_rows(size_t(g->n), 0);
    std::vector<const float*> q_dev(size_t(g->n), nullptr);
    for (int s = 0; s < g->n; s++) {
        mvdb_index* ix = g->shards[size_t(s)];
        mvdb_workspace* ws = g->ws[size_t(s)];
        CU_OK(cudaSetDevice(ix->device));
        if (on > g->out_cap || !g->D_dev[size_t(s)]) {
            cudaFree(g->D_dev[size_t(s)]);
            cudaFree(g->I_dev[size_t(s)]);
            g->D_dev[size_t(s)] = nullptr;
            g->I_dev[size_t(s)] = nullptr;
            CU_OK(cudaMalloc(&g->D_dev[size_t(s)], std::max(on, g->out_cap) * 4));
            CU_OK(cudaMalloc(&g->I_dev[size_t(s)], std::max(on, g->out_cap) * 8));
        }
        RC_OK(prepare_fused_scan(ix, ws, nq, k));
        const mvdb_mask* h = masks ? masks[s] : nullptr;
        const uint8_t* hm = (!h && host_masks) ? host_masks[s] : nullptr;
        if (h && h->ix != ix) return fail(MVDB_ERR_ARG, "mask handle %d belongs to another index", s);
        size_t words = 0, q_at = 0;
        if (hm) {
            mask_rows[size_t(s)] = std::min<uint64_t>(host_mask_rows ? host_mask_rows[s] : 0, ix->ntotal.load(std::memory_order_acquire));
            words = (mask_rows[size_t(s)] + 31) / 32;
            q_at = align_up(words * 4, 128) / 4;
        }
        // one transfer per shard: [filter words | pad | queries]
        RC_OK(grow_dev(&ws->mask_dev, &ws->mask_cap, q_at + qn));
        RC_OK(grow_pin(&ws->mask_pin, &ws->mask_pin_cap, q_at + qn));
        if (hm && words) {
            const size_t bytes = (mask_rows[size_t(s)] + 7) / 8;
            ws->mask_pin[words - 1] = 0;
            memcpy(ws->mask_pin, hm, bytes);
            if (mask_rows[size_t(s)] & 7) reinterpret_cast<uint8_t*>(ws->mask_pin)[bytes - 1] &= uint8_t((1u << (mask_rows[size_t(s)] & 7)) - 1u);
        }
        memcpy(ws->mask_pin + q_at, q, qn * 4);
        q_dev[size_t(s)] = reinterpret_cast<const float*>(ws->mask_dev + q_at);
        if (hm) mask_dev[size_t(s)] = ws->mask_dev;
        if (h) {
            mask_dev[size_t(s)] = h->dev;
            mask_rows[size_t(s)] = h->rows;
        }
    }
    if (on > g->out_cap) g->out_cap = on;
    // phase 2: copies and launches, device after device; nothing here waits for the GPU
    int rc = MVDB_OK;
    for (int s = 0; s < g->n && rc == MVDB_OK; s++) {
        mvdb_index* ix = g->shards[size_t(s)];
        mvdb_workspace* ws = g->ws[size_t(s)];
        cudaError_t e = cudaSetDevice(ix->device);
        const size_t q_at = size_t(reinterpret_cast<const uint32_t*>(q_dev[size_t(s)]) - ws->mask_dev);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ws->mask_dev, ws->mask_pin, (q_at + qn) * 4, cudaMemcpyHostToDevice, ws->stream);
        if (e != cudaSuccess) {
            rc = fail(MVDB_ERR_CUDA, "staging for shard %d failed: %s", s, cudaGetErrorString(e));
            break;
        }
        if (masks && masks[s]) rc = mask_wait(masks[s], ws->stream);
        if (rc == MVDB_OK)
            rc = run_search(ix, ws, q_dev[size_t(s)], nq, k, mask_dev[size_t(s)], mask_rows[size_t(s)], normalize_queries, 0,
                            g->D_dev[size_t(s)], g->I_dev[size_t(s)], ws->stream, g->xch[size_t(s)]);
    }
    // phase 3: the merged list is on every device; read it from shard 0 ([labels | distances], one sync)
    mvdb_workspace* ws0 = g->ws[0];
    cudaSetDevice(g->shards[0]->device);
    if (rc == MVDB_OK) {
        rc = grow_pin(&ws0->I_pin, &ws0->I_pin_cap, on + (on + 1) / 2);
        if (rc == MVDB_OK) {
            cudaError_t e = cudaMemcpyAsync(ws0->I_pin, g->I_dev[0], on * 8, cudaMemcpyDeviceToHost, ws0->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ws0->I_pin + on, g->D_dev[0], on * 4, cudaMemcpyDeviceToHost, ws0->stream);
            if (e != cudaSuccess) rc = fail(MVDB_ERR_CUDA, "result copy failed: %s", cudaGetErrorString(e));
        }
    }
    // always drain every shard's stream: a failed launch on one device leaves the others waiting for their timeout
    for (int s = 0; s < g->n; s++) {
        cudaSetDevice(g->shards[size_t(s)]->device);
        cudaError_t e = cudaStreamSynchronize(g->ws[size_t(s)]->stream);
        if (e != cudaSuccess && rc == MVDB_OK) rc = fail(MVDB_ERR_CUDA, "shard %d failed: %s", s, cudaGetErrorString(e));
    }
    if (rc != MVDB_OK) return rc;
    for (int s = 0; s < g->n; s++)
        if (*reinterpret_cast<volatile unsigned int*>(g->xch[size_t(s)]->status))
            return fail(MVDB_ERR_STATE, "shard %d gave up waiting for a peer's top-k (exchange timeout)", s);
    memcpy(I, ws0->I_pin, on * 8);
    memcpy(D, ws0->I_pin + on, on * 4);
    return MVDB_OK;
}

int mvdb_debug_gemm_scores(mvdb_index* ix, const float* q, int64_t nq, float* out) {
    ENTER(ix);
    if (!q || !out || nq <= 0) return fail(MVDB_ERR_ARG, "bad arguments");
    const uint64_t n = ix->ntotal.load(std::memory_order_acquire);
    if (n == 0 || n > 0xFFFFFFF0ull) return fail(MVDB_ERR_ARG, "index empty or too large");
    mvdb_workspace* ws = nullptr;
    RC_OK(pool_acquire(ix, &ws));
    struct Release {
        mvdb_index* ix;
        mvdb_workspace* ws;
        ~Release() { pool_release(ix, ws); }
    } rel{ix, ws};
    std::shared_lock<std::shared_mutex> mv(ix->move_mu);
    float *q_dev = nullptr, *o_dev = nullptr;
    CU_OK(cudaMalloc(&q_dev, size_t(nq) * ix->d * 4));
    cudaError_t e = cudaMalloc(&o_dev, size_t(nq) * n * 4);
    int rc = MVDB_OK;
    if (e != cudaSuccess) rc = fail(MVDB_ERR_OOM, "cudaMalloc failed");
    if (rc == MVDB_OK) {
        e = cudaMemcpyAsync(q_dev, q, size_t(nq) * ix->d * 4, cudaMemcpyHostToDevice, ws->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(o_dev, 0, size_t(nq) * n * 4, ws->stream);
        if (e != cudaSuccess) rc = fail(MVDB_ERR_CUDA, "copy failed: %s", cudaGetErrorString(e));
    }
    if (rc == MVDB_OK) rc = run_batched(ix, ws, q_dev, nq, 1, nullptr, uint32_t(n), 0, 0, nullptr, nullptr, ws->stream, 2, o_dev);
    if (rc == MVDB_OK) {
        e = cudaMemcpyAsync(out, o_dev, size_t(nq) * n * 4, cudaMemcpyDeviceToHost, ws->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ws->stream);
        if (e != cudaSuccess) rc = fail(MVDB_ERR_CUDA, "debug gemm failed: %s", cudaGetErrorString(e));
    }
    cudaFree(q_dev);
    cudaFree(o_dev);
    return rc;
}

int mvdb_debug_read_gemm_prof(mvdb_index* ix, uint64_t* out, int ctas) {
    ENTER(ix);
    if (!out || ctas <= 0 || ctas > 256 || !ix->gemm_prof_dev)
        return fail(MVDB_ERR_STATE, "GEMM profiling is off (set option \"gemm_prof\" = 1) or bad arguments");
    CU_OK(cudaDeviceSynchronize());
    CU_OK(cudaMemcpy(out, ix->gemm_prof_dev, size_t(ctas) * 8 * 8, cudaMemcpyDeviceToHost));
    return MVDB_OK;
}

int mvdb_debug_read_shadow_counters(mvdb_workspace* ws, uint32_t* out4) {
    if (!ws || !out4) return fail(MVDB_ERR_ARG, "null argument");
    if (!ws->ix) return fail(MVDB_ERR_STATE, "the workspace's index has been destroyed");
    if (!ws->i8_ctl && !ws->sv_ctl) return fail(MVDB_ERR_STATE, "this workspace has run neither an int8 shadow nor a survivor-list search");
    DeviceGuard guard(ws->ix->device);
    CU_OK(cudaDeviceSynchronize());
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (ws->i8_ctl) {
        I8Ctl c;
        CU_OK(cudaMemcpy(&c, ws->i8_ctl, sizeof c, cudaMemcpyDeviceToHost));
        out4[0] = c.last_cand;
        out4[1] = c.last_surv;
        out4[2] = c.overflow;
    }
    if (ws->sv_ctl) {
        SurvCtl c;
        CU_OK(cudaMemcpy(&c, ws->sv_ctl, sizeof c, cudaMemcpyDeviceToHost));
        out4[3] = c.last_count;
        out4[2] |= c.overflow << 1;
    }
    return MVDB_OK;
}

int mvdb_debug_read_trace(mvdb_index* ix, uint64_t* out16) {
    ENTER(ix);
    if (!out16 || !ix->trace_dev) return fail(MVDB_ERR_STATE, "tracing is off (set option \"trace\" = 1)");
    CU_OK(cudaDeviceSynchronize());
    CU_OK(cudaMemcpy(out16, ix->trace_dev, 16 * 8, cudaMemcpyDeviceToHost));
    return MVDB_OK;
}

int mvdb_normalize_L2(float* x, uint64_t n, int d, int device) {
    if (d <= 0) return fail(MVDB_ERR_ARG, "dimension %d out of range", d);
    if (n == 0) return MVDB_OK;
    if (!x) return fail(MVDB_ERR_ARG, "null rows");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(MVDB_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(MVDB_ERR_ARG, "device %d out of range", device);
    DeviceGuard guard(device);
    // the ingest kernel itself, writing to a second dense buffer (ld = d): a row normalised here is
    // bit-identical to the same row normalised by mvdb_index_add
    float *dev = nullptr, *out = nullptr;
    int* bits = nullptr;
    const size_t bytes = size_t(n) * d * 4;
    CU_OK(cudaMalloc(&dev, bytes));
    cudaError_t e = cudaMalloc(&out, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&bits, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(bits, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(dev, x, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        unsigned grid = unsigned(std::min<uint64_t>((n * 32 + 255) / 256, 148ull * 16));
        append_rows_kernel<false><<<grid, 256>>>(dev, out, n, d, int64_t(d), 1, 0, 0, 0, bits);
        LAUNCHED();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(x, out, bytes, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    cudaFree(out);
    cudaFree(bits);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? MVDB_ERR_OOM : MVDB_ERR_CUDA, "normalize_L2 failed: %s", cudaGetErrorString(e));
    return MVDB_OK;
}

int mvdb_merge_topk_device(int device, const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int64_t k,
                           float* D_out, int64_t* I_out, void* stream) {
    if (nparts <= 0 || nq < 0 || k <= 0) return fail(MVDB_ERR_ARG, "bad merge shape");
    if (nq == 0) return MVDB_OK;
    if (!D_parts || !I_parts || !D_out || !I_out) return fail(MVDB_ERR_ARG, "null buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MVDB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    uint64_t total = uint64_t(nparts) * uint64_t(k);
    uint64_t npad = 64;
    while (npad < total) npad <<= 1;
    if (npad > 16384) return fail(MVDB_ERR_ARG, "merge of %d x k=%lld exceeds 16384 candidates per query", nparts, (long long)k);
    if (npad * 8 > 48 * 1024)
        CU_OK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(npad * 8)));
    merge_topk_kernel<<<unsigned(nq), 256, npad * 8, static_cast<cudaStream_t>(stream)>>>(D_parts, I_parts, nparts, nq, k,
                                                                                         uint32_t(npad), D_out, I_out);
    LAUNCHED();
    CU_OK(cudaGetLastError());
    return MVDB_OK;
}

}  // extern "C"
