// gemm_tc.cuh -- large-batch path: Q[nq,d] x X[N,d]^T on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with a
// threshold-filter epilogue that never materialises the nq x N score matrix.
//
// Replaces faiss's exhaustive_inner_product_blas + Reservoir/Heap handlers (the
// nq >= 20 branch of IndexFlatIP.search; reference call site
// minivectordb/vector_database.py:497 -- the reference itself only ever sends
// one query, BASELINE.json config 3 is the batched extension).
//
// Operands are bf16 (a shadow copy of the fp32 matrix, rounded to nearest at
// ingest; queries converted per batch), accumulation is fp32 in TMEM.
//   * "bf16" mode returns the top-k of those scores (recall reported vs fp32);
//   * "exact" mode keeps every row whose bf16 score is within a RIGOROUS error
//     bound of the running k-th best (|q~.x~ - q.x| <= 2^-8 (1+2^-10) |q||x| for
//     round-to-nearest bf16 inputs, plus fp32 accumulation slack) and re-scores
//     the survivors in fp32 with the GEMV path's exact summation order, so
//     ids and distances are bit-identical to the single-query scan;
//   * "tf32" mode runs the same kernels with kind::tf32 over the fp32 matrix and the fp32
//     queries themselves (32 fp32 per 128-byte k-block, K = 8 per instruction; no shadow copy).
//
// Tile: 128 queries (UMMA M, one TMEM lane per query) x 256 rows (UMMA N, one
// TMEM column per row) x 64 bf16 of K per stage (= 128 B, SWIZZLE_128B).
// Warp roles (384 threads, 1 CTA/SM, persistent):
//   warp 0  TMA producer   cp.async.bulk.tensor.2d -> 4-stage smem ring
//   warp 1  MMA issuer     one elected lane, tcgen05.mma.cta_group::1.kind::f16; the next k-block's
//                          barrier is PEEKED inside the same asm block (umma_*_x4_peek) because any
//                          latency of this thread between two MMAs is a bubble in the tensor pipe
//   warp 2  TMEM allocator 512 columns = 2 accumulator stages x 256
//   warps 4-11 epilogue    tcgen05.ld 32x32b.x32: thread <-> query, columns <-> rows; two warps per
//                          TMEM lane quadrant, each taking 128 of the 256 columns
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <cfloat>

#include "aux_kernels.cuh"
#include "select.cuh"

namespace mvdb {

constexpr int kGemmBM = 128;      // queries per tile
constexpr int kGemmBN = 256;      // rows per tile
constexpr int kGemmBK = 64;       // bf16 per k-block (128 bytes)
constexpr int kGemmStages = 4;
constexpr int kGemmMaxStages = 8;  // single-CTA kernel with a short A tile (a_rows < 128): up to 6 stages
constexpr uint32_t kGemmABytes = kGemmBM * kGemmBK * 2;   // 16 KB
constexpr uint32_t kGemmBBytes = kGemmBN * kGemmBK * 2;   // 32 KB
constexpr uint32_t kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr uint32_t kGemmSmemBytes = kGemmStages * kGemmStageBytes + 256 + 1024;  // + barriers + alignment slack
constexpr uint32_t kTmemCols = 512;

struct GemmParams {
    int64_t nq;            // queries
    uint32_t row0, row1;   // SAMPLED rows [row0, row1) are scanned by this launch
    uint32_t n_valid;      // sampled rows >= n_valid do not exist
    uint32_t row_stride;   // sampled row i is index row i * row_stride (the tensor map and the
                           // bitmasks handed to this launch are already in sampled-row space)
    int d;
    const uint32_t* live;  // bitmasks over rows (nullptr = none)
    const uint32_t* mask;
    const uint32_t* const* qmask;   // [nq] per-query admissible bitmasks (nullptr array or nullptr entries = none)
    const uint32_t* qmask_words;    // [nq] 32-row words available behind qmask[q]; rows past them are not admissible
    // candidate output (threshold filter)
    const uint64_t* thr;   // [nq] threshold keys: a candidate must be > thr[q]
    uint64_t* cand;        // [nq][cand_cap] keys (bf16-score image << 32 | ~row)
    unsigned int* cand_cnt;  // [nq]
    uint32_t cand_cap;
    int l2_hint;           // 0 none, 1 keep X tiles (evict_last), 2 keep X and Q
    // debug: dense scores [nq][n_total] (nullptr in production)
    float* dense;
    int64_t dense_ld;
    // debug: per-CTA cycle counters [gridDim.x][8] (nullptr in production): 0 MMA thread total, 1 its wait
    // for operands (full), 2 its wait for a drained accumulator (tempty), 3 producer total, 4 producer wait
    // for a free stage (empty), 5 epilogue warp 4 total, 6 its wait for an accumulator (tfull)
    // 7 wall time of the MMA thread in ns (globaltimer)
    unsigned long long* prof;
    // single-CTA kernel only: query rows actually staged per k-block (32 / 64 / 128) and ring depth.
    // A batch of <= 32 queries stages 4 KB of A instead of 16 KB per k-block, which buys a 6-deep
    // instead of a 4-deep ring: the small-batch pass is HBM-bound and lives on bytes in flight.
    // (The MMA still reads 128 rows; rows past a_rows alias the B tile and feed accumulator lanes
    // that no epilogue thread reads.)
    uint32_t a_rows, stages;
    int debug;   // experiments, results are garbage: 1 = no operand loads (barriers only), 2 = no epilogue work, 4 = epilogue reads TMEM but skips the reduction,
                 // 8 = MMA thread never waits for operands (use with 1)
};

// mbar_wait that adds the cycles it blocked to `acc` when profiling
__device__ __forceinline__ void mbar_wait_prof(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
    if (on) {
        const long long a = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - a;
    } else {
        mbar_wait(bar, parity);
    }
}

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// One k-block (4 MMAs, 32 B of K each) with two barrier PEEKS folded in: mbarrier.test_wait
// (non-blocking) on `peek0` / `peek1` is issued BEFORE the MMAs and its predicate is only read
// AFTER them, so the issuing thread never sits between two tcgen05.mma waiting for shared memory.
// Returns bit 0 = peek0 complete, bit 1 = peek1 complete (a peek that is switched off reads as 0).
// Generated for: one CTA / bf16, CTA pair / bf16 (issued by the leader only), one CTA / tf32.
#define MVDB_DEFINE_UMMA_X4_PEEK(NAME, GROUP, KIND)                                                                      \
    __device__ __forceinline__ uint32_t NAME(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,          \
                                             uint32_t accumulate_first, uint64_t* peek0, uint32_t parity0, uint32_t do0, \
                                             uint64_t* peek1, uint32_t parity1, uint32_t do1) {                          \
        uint32_t tok;                                                                                                    \
        asm volatile(                                                                                                    \
            "{\n\t.reg .pred p0, p1, pa, pt, d0, d1;\n\t"                                                                \
            ".reg .b32 t0, t1;\n\t"                                                                                      \
            ".reg .b64 a, b;\n\t"                                                                                        \
            "setp.ne.b32 d0, %8, 0;\n\t"                                                                                 \
            "setp.ne.b32 d1, %11, 0;\n\t"                                                                                \
            "setp.ne.b32 p0, 0, 0;\n\t"                                                                                  \
            "setp.ne.b32 p1, 0, 0;\n\t"                                                                                  \
            "@d0 mbarrier.test_wait.parity.shared::cta.b64 p0, [%6], %7;\n\t"                                            \
            "@d1 mbarrier.test_wait.parity.shared::cta.b64 p1, [%9], %10;\n\t"                                           \
            "setp.ne.b32 pa, %5, 0;\n\t"                                                                                 \
            "setp.eq.b32 pt, 0, 0;\n\t"                                                                                  \
            "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%1], %2, %3, %4, pa;\n\t"                                  \
            "add.u64 a, %2, 2;\n\t"                                                                                      \
            "add.u64 b, %3, 2;\n\t"                                                                                      \
            "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%1], a, b, %4, pt;\n\t"                                    \
            "add.u64 a, %2, 4;\n\t"                                                                                      \
            "add.u64 b, %3, 4;\n\t"                                                                                      \
            "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%1], a, b, %4, pt;\n\t"                                    \
            "add.u64 a, %2, 6;\n\t"                                                                                      \
            "add.u64 b, %3, 6;\n\t"                                                                                      \
            "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%1], a, b, %4, pt;\n\t"                                    \
            "selp.u32 t0, 1, 0, p0;\n\t"                                                                                 \
            "selp.u32 t1, 2, 0, p1;\n\t"                                                                                 \
            "or.b32 %0, t0, t1;\n\t}"                                                                                    \
            : "=r"(tok)                                                                                                  \
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate_first), "r"(smem_u32(peek0)),            \
              "r"(parity0), "r"(do0), "r"(smem_u32(peek1)), "r"(parity1), "r"(do1)                                       \
            : "memory");                                                                                                 \
        return tok;                                                                                                      \
    }
MVDB_DEFINE_UMMA_X4_PEEK(umma_bf16_x4_peek, "1", "f16")
MVDB_DEFINE_UMMA_X4_PEEK(umma_bf16_x4_peek_2cta, "2", "f16")
MVDB_DEFINE_UMMA_X4_PEEK(umma_tf32_x4_peek, "1", "tf32")
#undef MVDB_DEFINE_UMMA_X4_PEEK
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives lane (base_lane + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile whose rows are 128 B
// (64 bf16) and stored with the 128-byte swizzle TMA produces: 8-row groups are
// 1024 B apart (SBO), LBO unused, descriptor version 1 (sm_100), layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);   // bits  0-13 start address >> 4
    d |= uint64_t(1024u >> 4) << 32;              // bits 32-45 stride byte offset >> 4
    d |= uint64_t(1) << 46;                       // bits 46-47 descriptor version
    d |= uint64_t(2) << 61;                       // bits 61-63 SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M x N.
__device__ __forceinline__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4)                 // c_format  = F32
           | (1u << 7)               // a_format  = BF16
           | (1u << 10)              // b_format  = BF16
           | (uint32_t(N >> 3) << 17)
           | (uint32_t(M >> 4) << 24);
}

// kind::tf32: A/B are fp32 words in shared memory of which the tensor core uses the upper 19 bits
// (sign, 8 exponent, 10 mantissa bits: truncation), K = 8 per instruction (32 B, like bf16's K = 16)
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4)                 // c_format  = F32
           | (2u << 7)               // a_format  = TF32
           | (2u << 10)              // b_format  = TF32
           | (uint32_t(N >> 3) << 17)
           | (uint32_t(M >> 4) << 24);
}
// elements of K in one 128-byte k-block
template <bool kTf32>
__device__ __forceinline__ constexpr int gemm_bk() { return kTf32 ? 32 : 64; }

// Epilogue of one 128 x 256 accumulator tile for the query owned by this thread (TMEM lane).
// t_lane: TMEM address of this warp's lane quadrant and accumulator stage; tile_row0: first
// (sampled) row of the tile; adm: admissible bits of the tile's 256 rows (common masks).
// g_lo .. g_lo+3: the four 32-column groups (half of the tile) this warp handles -- two warps
// share each TMEM lane quadrant and split the 256 columns between them.
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, uint32_t t_lane, int64_t q, uint32_t tile_row0,
                                                   const uint32_t (&adm)[8], int g_lo) {
    const bool q_ok = q < p.nq;
    const uint64_t thr = (q_ok && p.thr) ? p.thr[q] : kEmptyKey;
    // threshold as (score, ~row): an empty threshold admits every finite score
    const float thr_score = thr == kEmptyKey ? -INFINITY : key_score(thr);
    const uint32_t thr_low = thr == kEmptyKey ? 0xFFFFFFFFu : uint32_t(thr);
    const uint32_t* qm = (q_ok && p.qmask) ? p.qmask[q] : nullptr;   // this query's own filter
    const uint32_t qm_words = qm ? p.qmask_words[q] : 0u;
    if (p.dense) {
#pragma unroll 1
        for (int c0 = 32 * g_lo; c0 < 32 * g_lo + 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + uint32_t(c0), v);
            if (q_ok) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t row = tile_row0 + c0 + j;
                    if (row < p.n_valid) p.dense[q * p.dense_ld + row] = __uint_as_float(v[j]);
                }
            }
        }
    } else {
        // Pass 1: which of my query's 256 scores beat its threshold?  Pure register
        // work, branch-free, no memory operation inside a divergent region.  The
        // group loop is deliberately NOT unrolled: fully unrolled the epilogue is
        // ~200 KB of SASS and becomes instruction-fetch bound (measured: 10x slower).
        uint32_t hit[4];
        uint32_t total = 0;
#pragma unroll 1
        for (int g = g_lo; g < g_lo + 4; g++) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + uint32_t(32 * g), v);
            uint32_t m = 0;
            // Common case first: once the thresholds are warm almost no group holds a candidate,
            // so reduce the 32 scores to their maximum (16 FMNMX3) and compare ONCE.  The epilogue
            // warps share their schedulers with the MMA-issuing thread: every instruction saved
            // here is an issue slot it gets sooner (measured: the per-element test cost the
            // tensor pipe 17% at d = 1024 and 40% at d = 384).  fmaxf drops NaNs, which never pass.
            float sub[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                float t = __uint_as_float(v[8 * b]);
#pragma unroll
                for (int j = 1; j < 8; j++) t = fmaxf(t, __uint_as_float(v[8 * b + j]));
                sub[b] = t;
            }
            float mx = fmaxf(fmaxf(sub[0], sub[1]), fmaxf(sub[2], sub[3]));
            if (p.debug & 4) mx = __uint_as_float(v[0] & v[31]);   // experiment: TMEM reads only, no reduction
            if (q_ok && mx >= thr_score) {
                // ~row of column 0 of this group; column j is index row (tile_row0 + 32g + j) * stride
                const uint32_t low0 = 0xFFFFFFFFu - (tile_row0 + 32 * g) * p.row_stride;
                // rare path, still kept short: only the 8-score blocks whose own maximum passes are
                // examined score by score (hits are sparse, so that is nearly always one of the four)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (sub[b] >= thr_score) {
#pragma unroll
                        for (int j = 8 * b; j < 8 * b + 8; j++) {
                            // key(s,row) > thr  <=>  s > t, or s == t and ~row > thr_low  (float compares: NaN never
                            // passes, -0 == +0 exactly as the key's canonicalisation has it)
                            const float sc = __uint_as_float(v[j]);
                            const uint32_t pass = uint32_t(sc > thr_score) |
                                                  (uint32_t(sc == thr_score) & uint32_t((low0 - uint32_t(j) * p.row_stride) > thr_low));
                            m |= pass << j;
                        }
                    }
                }
                m &= adm[g];
                if (qm) {
                    const uint32_t w = (tile_row0 >> 5) + uint32_t(g);
                    m &= (w < qm_words) ? qm[w] : 0u;
                }
            }
            hit[g - g_lo] = m;
            total += __popc(m);
        }
        // One reservation per thread per tile, then (rarely) pass 2: re-read the groups
        // that had a hit (TMEM reads are cheap) and store the keys.
        uint32_t pos = 0;
        if (total) pos = atomicAdd(p.cand_cnt + q, total);
        if (__any_sync(0xFFFFFFFFu, total != 0u)) {
#pragma unroll 1
            for (int g = g_lo; g < g_lo + 4; g++) {
                uint32_t m = hit[g - g_lo];
                if (!__any_sync(0xFFFFFFFFu, m != 0u)) continue;
                uint32_t v[32];
                tmem_ld_32x32(t_lane + uint32_t(32 * g), v);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if ((m >> j) & 1u) {
                        if (pos < p.cand_cap)
                            p.cand[size_t(q) * p.cand_cap + pos] =
                                make_key(__uint_as_float(v[j]), (tile_row0 + 32 * g + j) * p.row_stride);
                        pos++;
                    }
                }
            }
        }
    }
}

struct GemmBarriers {
    uint64_t full[kGemmMaxStages];
    uint64_t empty[kGemmMaxStages];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
};

template <bool kTf32>
__global__ void __launch_bounds__(384, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
    constexpr int kBK = gemm_bk<kTf32>();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t n_stages = p.stages, a_bytes = p.a_rows * 128u, stage_bytes = a_bytes + kGemmBBytes;
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(smem + n_stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < n_stages; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&bars->tfull[a], 1);
            mbar_init(&bars->tempty[a], 8);   // one arrival per epilogue warp
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const uint32_t n_rows = p.row1 - p.row0;
    const uint32_t n_xt = (n_rows + kGemmBN - 1) / kGemmBN;
    const uint32_t n_qb = uint32_t((p.nq + kGemmBM - 1) / kGemmBM);
    const uint32_t n_kb = uint32_t((p.d + kBK - 1) / kBK);
    // tile t = xt * n_qb + qb; CTA c owns the contiguous range [t_lo, t_hi): consecutive
    // tiles share the X tile (L2 reuse) and small row chunks still fill the grid.
    const uint64_t n_tiles = uint64_t(n_xt) * n_qb;
    const uint32_t t_lo = uint32_t(n_tiles * blockIdx.x / gridDim.x);
    const uint32_t t_hi = uint32_t(n_tiles * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t keep = policy_evict_last();
            for (uint32_t t = t_lo; t < t_hi; t++) {
                const uint32_t xt = t / n_qb, qb = t % n_qb;
                    for (uint32_t kb = 0; kb < n_kb; kb++) {
                        mbar_wait(&bars->empty[stage], phase ^ 1u);
                        uint8_t* sA = smem + stage * stage_bytes;
                        uint8_t* sB = sA + a_bytes;
                        mbar_arrive_expect_tx(&bars->full[stage], stage_bytes);
                        if (p.l2_hint >= 2) tma_load_2d_hint(sA, &tmQ, int(kb * kBK), int(qb * kGemmBM), &bars->full[stage], keep);
                        else tma_load_2d(sA, &tmQ, int(kb * kBK), int(qb * kGemmBM), &bars->full[stage]);
                        if (p.l2_hint >= 1) tma_load_2d_hint(sB, &tmX, int(kb * kBK), int(p.row0 + xt * kGemmBN), &bars->full[stage], keep);
                        else tma_load_2d(sB, &tmX, int(kb * kBK), int(p.row0 + xt * kGemmBN), &bars->full[stage]);
                        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = kTf32 ? umma_idesc_tf32(kGemmBM, kGemmBN) : umma_idesc_bf16(kGemmBM, kGemmBN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            // barrier peeks folded into the MMA issue: see gemm_topk_kernel_mc
            uint32_t tok_full = 0, tok_acc = 0;
            for (uint32_t t = t_lo; t < t_hi; t++) {
                if (!tok_acc) mbar_wait(&bars->tempty[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kGemmBN;
                for (uint32_t kb = 0; kb < n_kb; kb++) {
                    if (!tok_full) mbar_wait(&bars->full[stage], phase);
                    const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
                    const uint64_t a_desc = umma_desc_sw128(a_addr);
                    const uint64_t b_desc = umma_desc_sw128(a_addr + a_bytes);
                    const uint32_t cur = stage;
                    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                    const bool last_kb = kb + 1 == n_kb;
                    const bool more = !(last_kb && t + 1 == t_hi);
                    const uint32_t tok = (kTf32 ? umma_tf32_x4_peek : umma_bf16_x4_peek)(d_tmem, a_desc, b_desc, idesc, kb != 0 ? 1u : 0u,
                                                           &bars->full[stage], phase, more ? 1u : 0u,
                                                           &bars->tempty[acc ^ 1u], acc_phase ^ (acc ^ 1u),
                                                           (more && last_kb) ? 1u : 0u);
                    tok_full = tok & 1u;
                    if (last_kb) tok_acc = tok >> 1;
                    umma_commit(&bars->empty[cur]);   // frees the smem slot when these MMAs retire
                }
                umma_commit(&bars->tfull[acc]);       // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        const uint32_t quad = uint32_t(warp & 3);            // the TMEM lane quadrant this warp may read (warp % 4)
        const int g_lo = warp >= 8 ? 4 : 0;                   // warps 4-7: columns 0-127, warps 8-11: columns 128-255
        uint32_t acc = 0, acc_phase = 0;
        uint32_t cur_xt = 0xFFFFFFFFu, tile_row0 = 0;
        uint32_t adm[8];
        for (uint32_t t = t_lo; t < t_hi; t++) {
            const uint32_t xt = t / n_qb, qb = t % n_qb;
            if (xt != cur_xt) {
                cur_xt = xt;
                tile_row0 = p.row0 + xt * kGemmBN;
                // admissible bits of the tile's 256 rows (8 words), same for every query
#pragma unroll
                for (int w = 0; w < 8; w++) {
                const uint32_t r = tile_row0 + 32 * w;        // tile_row0 is a multiple of 32 (row0 is, see host)
                uint32_t bits = 0xFFFFFFFFu;
                if (r >= p.n_valid) bits = 0;
                else {
                    if (p.n_valid - r < 32) bits = (1u << (p.n_valid - r)) - 1u;
                    if (p.mask) bits &= p.mask[r >> 5];
                    if (p.live) bits &= p.live[r >> 5];
                }
                    adm[w] = bits;
                }
            }
            {
                const int64_t q = int64_t(qb) * kGemmBM + quad * 32 + lane;
                mbar_wait(&bars->tfull[acc], acc_phase);
                tc_fence_after();
                gemm_epilogue_tile(p, tmem_base + ((quad * 32u) << 16) + acc * kGemmBN, q, tile_row0, adm, g_lo);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a CTA PAIR (cluster of 2, same TPC) computes one
// 256-query x 256-row tile.  Each CTA stages only ITS half of both operands -- 128
// queries (A) and 128 rows (B), 32 KB per stage instead of 48 KB -- and the leader
// CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256), which reads both
// CTAs' shared memory and writes a 128 x 256 accumulator into each CTA's TMEM.
// Per SM this halves the B-operand traffic from L2 and makes room for a 6-stage ring.
//
// Synchronisation:
//   full[s]      LEADER only: armed by the leader's producer with the bytes of BOTH CTAs' halves;
//                the peer's TMA loads complete their transaction bytes on the leader's barrier
//                directly (cp.async.bulk.tensor ... .cta_group::2 with the leader's barrier
//                address), so the MMA thread waits on ONE barrier and no relay is needed
//   empty[s]     each CTA: the leader's tcgen05.commit multicasts one arrival to both
//   tfull[a]     each CTA: accumulator a complete (multicast commit)
//   tempty[a]    leader only, count 16: the 8 epilogue warps of BOTH CTAs (peer: remote arrive)
// ---------------------------------------------------------------------------
constexpr int kGemm2Stages = 7;
constexpr uint32_t kGemm2HalfB = (kGemmBN / 2) * kGemmBK * 2;            // 16 KB: this CTA's 128 rows of the B tile
constexpr uint32_t kGemm2StageBytes = kGemmABytes + kGemm2HalfB;          // 32 KB
constexpr uint32_t kGemm2SmemBytes = kGemm2Stages * kGemm2StageBytes + 512 + 1024;
static_assert(kGemm2SmemBytes <= 232448, "2-CTA ring exceeds the 227 KB shared-memory opt-in");

struct Gemm2Barriers {
    uint64_t full[kGemm2Stages];
    uint64_t empty[kGemm2Stages];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
// shared::cluster address of `ptr`'s offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* ptr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(ptr)), "r"(rank));
    return ra;
}
// TMA load of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes complete
// on `bar_cluster_addr`, which may be a barrier of the peer CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst_smem, const CUtensorMap* map, int c0, int c1,
                                                uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one arrival on the barrier at this offset in EVERY CTA of `mask`, once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_topk_kernel_2cta(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Gemm2Barriers* bars = reinterpret_cast<Gemm2Barriers*>(smem + kGemm2Stages * kGemm2StageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemm2Stages; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&bars->tfull[a], 1);
            mbar_init(&bars->tempty[a], 16);  // 8 epilogue warps x 2 CTAs (waited on by the leader only)
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 2) tmem_alloc2(&bars->tmem_base, kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers exist before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const uint32_t n_rows = p.row1 - p.row0;
    const uint32_t n_xt = (n_rows + kGemmBN - 1) / kGemmBN;
    const uint32_t n_qb2 = uint32_t((p.nq + 2 * kGemmBM - 1) / (2 * kGemmBM));   // 256-query blocks
    const uint32_t n_kb = uint32_t((p.d + kGemmBK - 1) / kGemmBK);
    const uint64_t n_tiles = uint64_t(n_xt) * n_qb2;
    const uint32_t n_pairs = gridDim.x / 2, pair = blockIdx.x / 2;
    const uint32_t t_lo = uint32_t(n_tiles * pair / n_pairs);
    const uint32_t t_hi = uint32_t(n_tiles * (pair + 1) / n_pairs);

    if (warp == 0) {
        if (lane == 0) {   // TMA producer: this CTA's halves of A and B, signalled on the LEADER's barrier
            uint32_t stage = 0, phase = 0;
            const bool prof = p.prof != nullptr;
            long long w_empty = 0;
            const long long t_begin = clock64();
            for (uint32_t t = t_lo; t < t_hi; t++) {
                const uint32_t xt = t / n_qb2, qb2 = t % n_qb2;
                for (uint32_t kb = 0; kb < n_kb; kb++) {
                    mbar_wait_prof(&bars->empty[stage], phase ^ 1u, prof, w_empty);
                    uint8_t* sA = smem + stage * kGemm2StageBytes;
                    uint8_t* sB = sA + kGemmABytes;
                    const uint32_t full0 = mapa_u32(&bars->full[stage], 0);
                    if (p.debug & 1) {   // experiment: synchronisation only, no operand traffic
                        if (leader) mbar_arrive(&bars->full[stage]);
                        if (++stage == kGemm2Stages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    if (leader) mbar_arrive_expect_tx(&bars->full[stage], 2 * kGemm2StageBytes);
                    tma_load_2d_2sm(sA, &tmQ, int(kb * kGemmBK), int(qb2 * 2 * kGemmBM + rank * kGemmBM), full0);
                    tma_load_2d_2sm(sB, &tmX, int(kb * kGemmBK), int(p.row0 + xt * kGemmBN + rank * (kGemmBN / 2)), full0);
                    if (++stage == kGemm2Stages) { stage = 0; phase ^= 1u; }
                }
            }
            if (prof) {
                p.prof[blockIdx.x * 8 + 3] = (unsigned long long)(clock64() - t_begin);
                p.prof[blockIdx.x * 8 + 4] = (unsigned long long)w_empty;
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {   // the pair's single MMA issuer
            const uint32_t idesc = umma_idesc_bf16(2 * kGemmBM, kGemmBN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const bool prof = p.prof != nullptr;
            long long w_full = 0, w_tempty = 0;
            const long long t_begin = clock64();
            unsigned long long g_begin;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_begin));
            // barrier peeks folded into the MMA issue: see gemm_topk_kernel_mc
            uint32_t tok_full = 0, tok_acc = 0;
            for (uint32_t t = t_lo; t < t_hi; t++) {
                if (!tok_acc) mbar_wait_prof(&bars->tempty[acc], acc_phase ^ 1u, prof, w_tempty);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kGemmBN;
                for (uint32_t kb = 0; kb < n_kb; kb++) {
                    if (!tok_full && !(p.debug & 8)) mbar_wait_prof(&bars->full[stage], phase, prof, w_full);
                    const uint32_t a_addr = smem_u32(smem + stage * kGemm2StageBytes);
                    const uint64_t a_desc = umma_desc_sw128(a_addr);
                    const uint64_t b_desc = umma_desc_sw128(a_addr + kGemmABytes);
                    const uint32_t cur = stage;
                    if (++stage == kGemm2Stages) { stage = 0; phase ^= 1u; }
                    const bool last_kb = kb + 1 == n_kb;
                    const bool more = !(last_kb && t + 1 == t_hi);
                    const uint32_t tok = umma_bf16_x4_peek_2cta(d_tmem, a_desc, b_desc, idesc, kb != 0 ? 1u : 0u,
                                                                &bars->full[stage], phase, more ? 1u : 0u,
                                                                &bars->tempty[acc ^ 1u], acc_phase ^ (acc ^ 1u),
                                                                (more && last_kb) ? 1u : 0u);
                    tok_full = tok & 1u;
                    if (last_kb) tok_acc = tok >> 1;
                    umma_commit_2cta(&bars->empty[cur], 0x3);   // frees the stage in both CTAs
                }
                umma_commit_2cta(&bars->tfull[acc], 0x3);       // accumulator ready in both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (prof) {
                p.prof[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - t_begin);
                p.prof[blockIdx.x * 8 + 1] = (unsigned long long)w_full;
                p.prof[blockIdx.x * 8 + 2] = (unsigned long long)w_tempty;
                unsigned long long g_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
                p.prof[blockIdx.x * 8 + 7] = g_end - g_begin;
            }
        }
    } else if (warp >= 4) {
        const uint32_t quad = uint32_t(warp & 3);
        const int g_lo = warp >= 8 ? 4 : 0;
        uint32_t acc = 0, acc_phase = 0;
        uint32_t cur_xt = 0xFFFFFFFFu, tile_row0 = 0;
        uint32_t adm[8];
        const bool prof = p.prof != nullptr;
        long long w_tfull = 0;
        const long long t_begin = clock64();
        for (uint32_t t = t_lo; t < t_hi; t++) {
            const uint32_t xt = t / n_qb2, qb2 = t % n_qb2;
            if (xt != cur_xt) {
                cur_xt = xt;
                tile_row0 = p.row0 + xt * kGemmBN;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint32_t r = tile_row0 + 32 * w;
                    uint32_t bits = 0xFFFFFFFFu;
                    if (r >= p.n_valid) bits = 0;
                    else {
                        if (p.n_valid - r < 32) bits = (1u << (p.n_valid - r)) - 1u;
                        if (p.mask) bits &= p.mask[r >> 5];
                        if (p.live) bits &= p.live[r >> 5];
                    }
                    adm[w] = bits;
                }
            }
            const int64_t q = int64_t(qb2) * 2 * kGemmBM + rank * kGemmBM + quad * 32 + lane;
            mbar_wait_prof(&bars->tfull[acc], acc_phase, prof, w_tfull);
            tc_fence_after();
            if (!(p.debug & 2)) gemm_epilogue_tile(p, tmem_base + ((quad * 32u) << 16) + acc * kGemmBN, q, tile_row0, adm, g_lo);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&bars->tempty[acc]);
                else mbar_arrive_remote(&bars->tempty[acc], 0);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (prof && warp == 4 && lane == 0) {
            p.prof[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - t_begin);
            p.prof[blockIdx.x * 8 + 6] = (unsigned long long)w_tfull;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody leaves while the peer may still touch its shared memory / barriers
    if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// Multicast variant: a cluster of 2 CTAs works on the SAME X tile with two
// different query blocks.  Each CTA fetches half of the X tile (128 rows) and TMA
// multicasts it into both CTAs' shared memory, so every L2 read of X feeds two SMs:
// L2 output per SM per k-block drops from 48 KB to 32 KB.  MMAs stay cta_group::1
// (each CTA owns its 128 x 256 tile); the only coupling is the `empty` barrier, which
// needs BOTH CTAs' MMA commits before either producer may overwrite the stage.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d_mc(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                               uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// CS = cluster size (2 or 4): CS query blocks share one X tile; each CTA fetches 1/CS of it.
template <int CS, bool kTf32>
__global__ void __launch_bounds__(384, 1)
gemm_topk_kernel_mc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmXh, const GemmParams p) {
    constexpr int kBK = gemm_bk<kTf32>();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(smem + kGemmStages * kGemmStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], CS);   // every CTA's MMA commit
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&bars->tfull[a], 1);
            mbar_init(&bars->tempty[a], 8);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmXh);
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const uint32_t n_rows = p.row1 - p.row0;
    const uint32_t n_xt = (n_rows + kGemmBN - 1) / kGemmBN;
    const uint32_t n_qb2 = uint32_t((p.nq + CS * kGemmBM - 1) / (CS * kGemmBM));   // groups of CS query blocks
    const uint32_t n_kb = uint32_t((p.d + kBK - 1) / kBK);
    const uint64_t n_tiles = uint64_t(n_xt) * n_qb2;
    const uint32_t n_pairs = gridDim.x / CS, pair = blockIdx.x / CS;
    const uint32_t t_lo = uint32_t(n_tiles * pair / n_pairs);
    const uint32_t t_hi = uint32_t(n_tiles * (pair + 1) / n_pairs);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const bool prof = p.prof != nullptr;
            long long w_empty = 0;
            const long long t_begin = clock64();
            for (uint32_t t = t_lo; t < t_hi; t++) {
                const uint32_t xt = t / n_qb2, qb = (t % n_qb2) * CS + rank;
                for (uint32_t kb = 0; kb < n_kb; kb++) {
                    mbar_wait_prof(&bars->empty[stage], phase ^ 1u, prof, w_empty);   // freed by BOTH CTAs
                    uint8_t* sA = smem + stage * kGemmStageBytes;
                    uint8_t* sB = sA + kGemmABytes;
                    if (p.debug & 1) {   // experiment: synchronisation only, no operand traffic
                        mbar_arrive(&bars->full[stage]);
                        if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&bars->full[stage], kGemmStageBytes);   // A + my half of B + the peer's half
                    tma_load_2d(sA, &tmQ, int(kb * kBK), int(qb * kGemmBM), &bars->full[stage]);
                    tma_load_2d_mc(sB + rank * (kGemmBBytes / CS), &tmXh, int(kb * kBK),
                                   int(p.row0 + xt * kGemmBN + rank * (kGemmBN / CS)), &bars->full[stage], uint16_t((1u << CS) - 1u));
                    if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
                }
            }
            if (prof) {
                p.prof[blockIdx.x * 8 + 3] = (unsigned long long)(clock64() - t_begin);
                p.prof[blockIdx.x * 8 + 4] = (unsigned long long)w_empty;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = kTf32 ? umma_idesc_tf32(kGemmBM, kGemmBN) : umma_idesc_bf16(kGemmBM, kGemmBN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const bool prof = p.prof != nullptr;
            long long w_full = 0, w_tempty = 0;
            const long long t_begin = clock64();
            unsigned long long g_begin;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_begin));
            // Any latency this thread spends between two tcgen05.mma is a bubble in the tensor pipe
            // (measured: a try_wait on an ALREADY complete barrier at every k-block boundary cost
            // ~90 of every 600 cycles; the pipe reaches 100% with the wait removed).  So the barrier
            // of the NEXT k-block (and, at a tile boundary, the next tile's drained accumulator) is
            // only PEEKED with a non-blocking test_wait issued ahead of the current k-block's MMAs
            // and read after them (umma_bf16_x4_peek); the blocking loop runs only if a peek failed.
            static_assert(kGemmBK == 64, "the x4_peek helpers issue exactly four 32-byte K steps per 128-byte k-block");
            uint32_t tok_full = 0, tok_acc = 0;
            for (uint32_t t = t_lo; t < t_hi; t++) {
                if (!tok_acc) mbar_wait_prof(&bars->tempty[acc], acc_phase ^ 1u, prof, w_tempty);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kGemmBN;
                for (uint32_t kb = 0; kb < n_kb; kb++) {
                    if (!tok_full && !(p.debug & 8)) mbar_wait_prof(&bars->full[stage], phase, prof, w_full);
                    const uint32_t a_addr = smem_u32(smem + stage * kGemmStageBytes);
                    const uint64_t a_desc = umma_desc_sw128(a_addr);
                    const uint64_t b_desc = umma_desc_sw128(a_addr + kGemmABytes);
                    const uint32_t cur = stage;
                    if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
                    const bool last_kb = kb + 1 == n_kb;
                    const bool more = !(last_kb && t + 1 == t_hi);
                    const uint32_t tok = (kTf32 ? umma_tf32_x4_peek : umma_bf16_x4_peek)(d_tmem, a_desc, b_desc, idesc, kb != 0 ? 1u : 0u,
                                                           &bars->full[stage], phase, more ? 1u : 0u,
                                                           &bars->tempty[acc ^ 1u], acc_phase ^ (acc ^ 1u),
                                                           (more && last_kb) ? 1u : 0u);
                    tok_full = tok & 1u;
                    if (last_kb) tok_acc = tok >> 1;
                    umma_commit_mc(&bars->empty[cur], uint16_t((1u << CS) - 1u));   // one arrival in each CTA's empty[cur]
                }
                umma_commit(&bars->tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (prof) {
                p.prof[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - t_begin);
                p.prof[blockIdx.x * 8 + 1] = (unsigned long long)w_full;
                p.prof[blockIdx.x * 8 + 2] = (unsigned long long)w_tempty;
                unsigned long long g_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
                p.prof[blockIdx.x * 8 + 7] = g_end - g_begin;
            }
        }
    } else if (warp >= 4) {
        const uint32_t quad = uint32_t(warp & 3);
        const int g_lo = warp >= 8 ? 4 : 0;
        uint32_t acc = 0, acc_phase = 0;
        uint32_t cur_xt = 0xFFFFFFFFu, tile_row0 = 0;
        uint32_t adm[8];
        const bool prof = p.prof != nullptr;
        long long w_tfull = 0;
        const long long t_begin = clock64();
        for (uint32_t t = t_lo; t < t_hi; t++) {
            const uint32_t xt = t / n_qb2, qb = (t % n_qb2) * CS + rank;
            if (xt != cur_xt) {
                cur_xt = xt;
                tile_row0 = p.row0 + xt * kGemmBN;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint32_t r = tile_row0 + 32 * w;
                    uint32_t bits = 0xFFFFFFFFu;
                    if (r >= p.n_valid) bits = 0;
                    else {
                        if (p.n_valid - r < 32) bits = (1u << (p.n_valid - r)) - 1u;
                        if (p.mask) bits &= p.mask[r >> 5];
                        if (p.live) bits &= p.live[r >> 5];
                    }
                    adm[w] = bits;
                }
            }
            const int64_t q = int64_t(qb) * kGemmBM + quad * 32 + lane;
            mbar_wait_prof(&bars->tfull[acc], acc_phase, prof, w_tfull);
            tc_fence_after();
            if (!(p.debug & 2)) gemm_epilogue_tile(p, tmem_base + ((quad * 32u) << 16) + acc * kGemmBN, q, tile_row0, adm, g_lo);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (prof && warp == 4 && lane == 0) {
            p.prof[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - t_begin);
            p.prof[blockIdx.x * 8 + 6] = (unsigned long long)w_tfull;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// helpers around the GEMM
// ---------------------------------------------------------------------------
// fp32 rows (pitch ld_in floats) -> bf16 rows (pitch ld_out), round to nearest, zero padding
__global__ void __launch_bounds__(256) to_bf16_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                           uint64_t n, int d, int64_t ld_in, int64_t ld_out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const float* src = in + r * ld_in;
        __nv_bfloat16* dst = out + r * ld_out;
        for (int c = lane; c < ld_out; c += kWarp) dst[c] = __float2bfloat16_rn(c < d ? src[c] : 0.f);
    }
}

// Queries: dense [nq][d] -> padded fp32 [nq][ld] (optionally L2-normalised with
// EXACTLY the arithmetic of the scan's load_query_regs, so that re-scored
// distances are bit-identical to the single-query path) + their L2 norms.
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ q, float* __restrict__ qn,
                                                           float* __restrict__ qnorm, int64_t nq, int d, int64_t ld,
                                                           int normalize) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nq) return;
    const float* src = q + warp * d;
    float* dst = qn + warp * ld;
    const int ld4 = int(ld >> 2);
    float nr = 0.f;
    for (int c = lane; c < ld4; c += kWarp) {   // same chunk order as load_query_regs (j ascending)
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int b = 4 * c;
        if (b + 0 < d) v.x = src[b + 0];
        if (b + 1 < d) v.y = src[b + 1];
        if (b + 2 < d) v.z = src[b + 2];
        if (b + 3 < d) v.w = src[b + 3];
        nr = dot4(v, v, nr);
        reinterpret_cast<float4*>(dst)[c] = v;
    }
    nr = warp_allsum(nr);
    float norm = sqrtf(nr);
    if (normalize && nr > 0.f) {
        const float inv = renorm_scale(nr);
        __syncwarp();
        for (int c = lane; c < ld4; c += kWarp) {
            float4 v = reinterpret_cast<float4*>(dst)[c];
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
            reinterpret_cast<float4*>(dst)[c] = v;
        }
        norm = 1.0f;
    }
    if (lane == 0) qnorm[warp] = norm;
}

// Bitmask in sampled-row space: bit i of dst = AND over the given source masks of bit (i * stride).
// srcs: up to two common masks (live, filter; nullptr = all ones).  One thread per output word.
__global__ void sample_mask_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t n_src_rows,
                                   uint32_t stride, uint32_t* __restrict__ dst, uint32_t m_rows) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (m_rows + 31) / 32) return;
    uint32_t out = 0;
    for (int j = 0; j < 32; j++) {
        const uint64_t i = uint64_t(w) * 32 + j;
        if (i >= m_rows) break;
        const uint64_t r = i * stride;
        if (r >= n_src_rows) break;
        uint32_t bit = 1u;
        if (a) bit &= (a[r >> 5] >> (r & 31)) & 1u;
        if (b) bit &= (b[r >> 5] >> (r & 31)) & 1u;
        out |= bit << j;
    }
    dst[w] = out;
}
// Same for per-query masks: query q's sampled mask goes to dst + q * words_per_query (all zero
// and pointer left null when the query has no filter).  grid.y = query.
__global__ void sample_qmask_kernel(const uint32_t* const* __restrict__ src, const uint32_t* __restrict__ src_words,
                                    uint32_t stride, uint32_t* __restrict__ dst, uint32_t m_rows, uint32_t words_per_query) {
    const uint32_t q = blockIdx.y;
    const uint32_t* s = src[q];
    if (!s) return;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (m_rows + 31) / 32) return;
    const uint64_t src_rows = uint64_t(src_words[q]) * 32;
    uint32_t out = 0;
    for (int j = 0; j < 32; j++) {
        const uint64_t i = uint64_t(w) * 32 + j;
        if (i >= m_rows) break;
        const uint64_t r = i * stride;
        if (r >= src_rows) break;
        out |= ((s[r >> 5] >> (r & 31)) & 1u) << j;
    }
    dst[size_t(q) * words_per_query + w] = out;
}
// pointer / length tables for a level's sampled per-query masks
__global__ void qmask_table_kernel(const uint32_t* const* __restrict__ src, uint32_t* dst_base, uint32_t words_per_query,
                                   const uint32_t** out_ptr, uint32_t* out_words, int64_t nq) {
    int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    out_ptr[q] = src[q] ? dst_base + size_t(q) * words_per_query : nullptr;
    out_words[q] = src[q] ? words_per_query : 0u;
}
// a new sampling level starts from an empty candidate list but keeps the thresholds
__global__ void reset_counts_kernel(unsigned int* cnt, int64_t nq) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < nq) cnt[i] = 0;
}

__global__ void init_batch_state_kernel(uint64_t* thr, unsigned int* cnt, unsigned int* overflow, int64_t nq) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < nq) {
        thr[i] = kEmptyKey;
        cnt[i] = 0;
        overflow[i] = 0;
    }
}

// Per query (one CTA): sort the candidate list best-first, trim it, refresh the
// threshold.  exact_slack == nullptr: keep the best k, threshold = k-th key.
// exact mode: threshold = (k-th bf16 score - slack[q]) and EVERYTHING above it is
// kept (the rigorous superset of the true fp32 top-k).  qnorm == nullptr selects
// the plain mode; otherwise slack = slack_unit * qnorm[q].
__global__ void __launch_bounds__(256) cand_update_kernel(uint64_t* cand, unsigned int* cnt, uint64_t* thr,
                                                          unsigned int* overflow, uint32_t cap, int k,
                                                          const float* qnorm, float slack_unit) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem);
    __shared__ unsigned int keep_s;
    const int64_t q = blockIdx.x;
    unsigned int n = cnt[q];
    if (n > cap) {
        if (threadIdx.x == 0) overflow[q] = 1u;
        n = cap;
    }
    if (n == 0) return;
    uint32_t npad = 64;
    while (npad < n) npad <<= 1;
    uint64_t* src = cand + size_t(q) * cap;
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) a[i] = i < n ? src[i] : kEmptyKey;
    __syncthreads();
    bitonic_sort_desc(a, int(npad), int(threadIdx.x), int(blockDim.x), BlockSyncer());
    uint64_t new_thr = kEmptyKey;
    unsigned int keep = n;
    if (n >= unsigned(k)) {
        const uint64_t kth = a[k - 1];
        if (qnorm == nullptr) {
            new_thr = kth;
            keep = unsigned(k);
        } else {
            const float t = key_score(kth) - slack_unit * qnorm[q];
            new_thr = make_key(t, 0xFFFFFFFFu);   // lowest key with score t: every score >= t passes
            if (threadIdx.x == 0) keep_s = 0;
            __syncthreads();
            unsigned int mine = 0;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mine += (a[i] > new_thr) ? 1u : 0u;
            if (mine) atomicAdd(&keep_s, mine);
            __syncthreads();
            keep = keep_s;
        }
    }
    for (uint32_t i = threadIdx.x; i < keep; i += blockDim.x) src[i] = a[i];
    if (threadIdx.x == 0) {
        cnt[q] = keep;
        // a threshold, once established (by a coarser sampling level or an earlier piece), is a
        // valid lower bound for good: a piece that yields fewer than k hits must not lower it
        if (new_thr > thr[q]) thr[q] = new_thr;
    }
}

// exact fp32 re-scoring of every surviving candidate, with the scan's summation
// order (per-lane chunk order + xor butterfly 16,8,4,2,1 == reduce8's tree).
__global__ void __launch_bounds__(256) rescore_kernel(const float* __restrict__ x, int ld4, const float* __restrict__ qn,
                                                      uint64_t* cand, const unsigned int* cnt, uint32_t cap) {
    extern __shared__ __align__(16) uint8_t smem[];
    float4* qs = reinterpret_cast<float4*>(smem);
    const int64_t q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const unsigned int n = min(cnt[q], cap);
    if (n == 0) return;
    const float4* qsrc = reinterpret_cast<const float4*>(qn) + q * ld4;
    for (int c = threadIdx.x; c < ld4; c += blockDim.x) qs[c] = qsrc[c];
    __syncthreads();
    uint64_t* list = cand + size_t(q) * cap;
    for (unsigned int i = warp; i < n; i += nw) {
        const uint32_t row = key_row(list[i]);
        const float4* xr = reinterpret_cast<const float4*>(x) + size_t(row) * ld4;
        float acc = 0.f;
        for (int c = lane; c < ld4; c += kWarp) acc = dot4(ldg_stream(xr + c), qs[c], acc);
        acc = warp_allsum(acc);
        if (lane == 0) list[i] = make_key(acc, row);
    }
}

// first k keys of every (sorted) candidate list -> (D, I)
__global__ void batch_results_kernel(const uint64_t* cand, const unsigned int* cnt, uint32_t cap, int k, int64_t nq,
                                     int64_t label_offset, float* D, int64_t* I) {
    int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= nq * k) return;
    const int64_t q = t / k;
    const int i = int(t % k);
    const uint64_t key = (unsigned(i) < min(cnt[q], cap)) ? cand[size_t(q) * cap + i] : kEmptyKey;
    if (key == kEmptyKey) {
        D[t] = -FLT_MAX;
        I[t] = -1;
    } else {
        D[t] = key_score(key);
        I[t] = int64_t(key_row(key)) + label_offset;
    }
}

}  // namespace mvdb
