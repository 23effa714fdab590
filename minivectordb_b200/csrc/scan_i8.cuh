// scan_i8.cuh -- opt-in low-precision scan mode ("scan_shadow" = 1): single-query search over an int8
// SHADOW of the matrix, exact fp32 re-scoring of a rigorous candidate superset.
//
// Why: the fp32 scan already streams the matrix at the HBM roofline, so the only lever left for the
// latency of ONE query is bytes.  The reference's own low-precision precedent is its usearch variant
// (ref minivectordb/sharded_vector_database_usearch.py:621-627, int8 vectors); unlike that variant this
// mode returns EXACTLY what the fp32 scan returns (ids and distances bit-identical, tested), because
// the int8 pass only selects a superset of the true top-k and the survivors are re-scored from the fp32
// master rows with the scan's own summation order.
//
// Shadow record of row r (stride rec_bytes = ld8 + 16, ld8 = d rounded up to 16):
//     [ld8 x int8  xq = rint(x / s_r)] [f32 s_r = max|x| / 127] [f32 e_r >= ||x - s_r xq||_2] [8 B pad]
// Query: two int8 planes q ~ q^ = s1 qq1 + s2 qq2 (the second plane quantises the residual of the
// first, so the query-side error rq = ||q - q^||_2 is ~1e-5 and costs only a second dp4a).
// For every row (Cauchy-Schwarz, all norms computed, not estimated):
//     | q.x - s_r (s1 dot1 + s2 dot2) |  <=  ||q^|| e_r + rq ||x||  =: B_r            (exact arithmetic)
// plus eps for the fp32 rounding of the scan's own score.  L_r = approx - B_r - eps is a lower bound of
// the fp32 scan's score of row r, U_r = approx + B_r + eps an upper bound.
// Threshold: every consumer warp publishes the best L it has seen (one word per warp, monotone); the
// k-th largest of those words is a lower bound of the final k-th best score (k distinct rows reach it),
// so a row with U_r below it can never enter the top-k.  Rows that pass are rare (a few hundred per
// search): the warp that found one re-scores it at once from the fp32 master row and appends the exact
// key to the survivor list; the last CTA to finish sorts the survivors and writes (D, I) -- one launch.
// HBM traffic: N (d + 16) bytes instead of N d 4.
#pragma once
#include "scan.cuh"

namespace mvdb {

constexpr int kI8TileRows = 32;      // one word of the bitmasks
constexpr int kI8MetaBytes = 16;
constexpr int kI8MaxJ = 4;           // 16-byte chunks per lane: d <= 2048
constexpr uint32_t kI8DynChunk = 4;  // tiles per claim of the counter-claimed tail
constexpr uint32_t kI8SurvCap = 4096;

struct I8Ctl {                // zero before the first search; every search leaves it clean for the next one
    unsigned int cand_cnt;    // rows the int8 pass could not rule out (statistics)
    unsigned int surv_cnt;    // of those, rows whose exact score reached the threshold: the survivor list
    unsigned int ticket;      // CTAs done
    unsigned int tile_ctr;    // counter-claimed tail of the tile schedule
    unsigned int overflow;    // 1: a list overflowed -> the conditional fp32 scan behind this search answers instead
    unsigned int last_cand;   // counters of the search that just finished (test hook)
    unsigned int last_surv;
    unsigned int pad[1];
};

struct I8Params {
    const uint8_t* x8;        // shadow records
    const float* x;           // fp32 master matrix (re-scoring)
    const float* q;           // [d] the query as given
    int normalize_q;
    const uint32_t* live;
    const uint32_t* mask;
    uint64_t* surv;           // [kI8SurvCap] exact keys of the re-scored candidates
    unsigned int* ovf_host;   // pinned host word raised on overflow, or nullptr (then only ctl->overflow is)
    unsigned int* best;       // [nbest] ordered images of the per-warp best lower bounds
    I8Ctl* ctl;
    float* outD;
    int64_t* outI;
    int64_t label_offset;
    uint32_t n, nbest, rec_bytes, stage_bytes, stage_off, q_off;
    uint32_t static_iters;    // iterations of every CTA served by the static round-robin (tile = cta + it * grid) ...
    uint32_t dyn_tile0;       // ... the tiles from here on are claimed from ctl->tile_ctr, one at a time (= one mask word)
    int d, ld4, ld8, k, stages;
    float max_norm;           // largest ||x_r|| stored in the index
    uint32_t dep_inputs;      // q and mask are written by the grid this launch programmatically depends on (see ScanParams)
    const XchgDev* xchg;      // fused cross-GPU exchange (nullptr = single GPU): the last CTA sends this shard's k best
    uint64_t xchg_seq;        // to every rank and merges, exactly as the fp32 scan's tail does
};

__device__ __forceinline__ int reduce8i(const int (&a)[8], int lane) {
    const unsigned full = 0xFFFFFFFFu;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    int b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int mine = h16 ? a[i + 4] : a[i];
        int give = h16 ? a[i] : a[i + 4];
        b[i] = mine + __shfl_xor_sync(full, give, 16);
    }
    int c[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int mine = h8 ? b[i + 2] : b[i];
        int give = h8 ? b[i] : b[i + 2];
        c[i] = mine + __shfl_xor_sync(full, give, 8);
    }
    int mine = h4 ? c[1] : c[0];
    int give = h4 ? c[0] : c[1];
    int s = mine + __shfl_xor_sync(full, give, 4);
    s += __shfl_xor_sync(full, s, 2);
    s += __shfl_xor_sync(full, s, 1);
    return s;
}

// ---------------------------------------------------------------------------
// fp32 rows -> int8 shadow records (one warp per row, built lazily like the bf16 shadow)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_i8_rows_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, uint64_t n,
                                                         int ld4, int ld8, uint32_t rec_bytes) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const int w8 = ld8 >> 2;   // 32-bit words of int8 per record
    for (uint64_t r = warp; r < n; r += nwarps) {
        const float4* src = reinterpret_cast<const float4*>(in) + r * uint64_t(ld4);
        uint8_t* rec = out + r * uint64_t(rec_bytes);
        float mx = 0.f;
        for (int c = lane; c < ld4; c += kWarp) {
            const float4 v = src[c];
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        const float s = mx * (1.0f / 127.0f);
        const float inv = s > 0.f ? 1.0f / s : 0.f;
        float e2 = 0.f;
        for (int c = lane; c < w8; c += kWarp) {
            uint32_t word = 0;
            if (c < ld4) {
                const float4 v = src[c];
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    int qv = __float2int_rn(f[t] * inv);
                    qv = max(-127, min(127, qv));
                    const float res = fmaf(-s, float(qv), f[t]);   // exact residual of this element (one rounding)
                    e2 = fmaf(res, res, e2);
                    word |= uint32_t(uint8_t(int8_t(qv))) << (8 * t);
                }
            }
            reinterpret_cast<uint32_t*>(rec)[c] = word;
        }
        e2 = warp_allsum(e2);
        if (lane == 0) {
            // rounded UP: the sum of squares and the square root each lose at most a few ulp
            float e = sqrtf(e2) * 1.00001f + 1e-30f;
            reinterpret_cast<float*>(rec + ld8)[0] = s;
            reinterpret_cast<float*>(rec + ld8)[1] = e;
            reinterpret_cast<float*>(rec + ld8)[2] = 0.f;
            reinterpret_cast<float*>(rec + ld8)[3] = 0.f;
        }
    }
}

// Scale that faiss.normalize_L2 would apply to the query, computed with EXACTLY the arithmetic of the fp32
// scan's load_query_regs (per-lane float4 chunks lane + 32 j in ascending j, xor-butterfly warp sum), so that
// the normalised query -- and with it every re-scored distance -- is bit-identical to the fp32 scan's.
// Returns 1 when no scaling applies.  q: dense [d] floats (any alignment).
__device__ __forceinline__ float4 i8_load_q4(const float* q, int d, int c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int b = 4 * c;
    if (b + 3 < d && (reinterpret_cast<uintptr_t>(q) & 15) == 0) return reinterpret_cast<const float4*>(q)[c];
    if (b + 0 < d) v.x = q[b + 0];
    if (b + 1 < d) v.y = q[b + 1];
    if (b + 2 < d) v.z = q[b + 2];
    if (b + 3 < d) v.w = q[b + 3];
    return v;
}
__device__ __forceinline__ float i8_query_scale(const float* q, int d, int ld4, int normalize, int lane, bool* scaled) {
    float nr = 0.f;
    for (int c = lane; c < ld4; c += kWarp) {
        const float4 v = i8_load_q4(q, d, c);
        nr = dot4(v, v, nr);
    }
    nr = warp_allsum(nr);
    *scaled = normalize && nr > 0.f;
    return *scaled ? renorm_scale(nr) : 1.0f;
}

// shared-memory header of the int8 scan
struct I8Header {
    uint64_t full[16];
    uint64_t empty[16];
    unsigned int thr;       // CTA-wide copy of the threshold (ordered image), only ever raised
    unsigned int refreshes; // how many times any warp of the CTA has recomputed it
    int last_flag;
    unsigned int tail_ns, tail_nc;   // last CTA: survivor / candidate counts (read once, by the thread that resets them)
    int cnts4[4];           // scratch of block_select_topk (tail)
    uint32_t tile_of[16];   // counter-claimed tiles: tile held by ring stage s (kNoTile = stop) ...
    uint32_t adm_of[16];    // ... and its admissible word (mask & live)
};

// ---------------------------------------------------------------------------
// the scan: warp 0 = TMA producer (one bulk copy of 32 records per tile), warps 1.. = consumers
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(288, 1) scan_i8_kernel(const I8Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    I8Header* hdr = reinterpret_cast<I8Header*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - 1;
    const int cw = warp - 1;
    const uint32_t G = gridDim.x;
    const uint32_t T = (p.n + kI8TileRows - 1) / kI8TileRows;
    const uint32_t iters = (T > blockIdx.x) ? (T - blockIdx.x + G - 1) / G : 0;
    const int S = p.stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&hdr->full[s], 1);
            mbar_init(&hdr->empty[s], 1);
        }
        if (blockIdx.x == 0) p.ctl->overflow = 0u;   // the previous search's conditional fallback has completed (stream order)
        hdr->thr = 0u;
        hdr->refreshes = 0u;
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            // Same two-phase schedule as the fp32 scan: a static round-robin first (no atomics, no start-up
            // latency), then the last share of the tiles claimed from a global counter so that the SMs finish
            // together.  Claims (and the claimed tile's mask word) run two ahead of their use.
            const uint64_t pol = policy_evict_first();
            int s = 0;
            uint32_t ph = 0;
            const uint32_t n_static = min(iters, p.static_iters);
            // claims are chunks of kI8DynChunk consecutive tiles (an atomic's latency is ~2 tiles of streaming)
            const uint32_t n_dyn = T > p.dyn_tile0 ? (T - p.dyn_tile0 + kI8DynChunk - 1) / kI8DynChunk : 0u;
            uint32_t c0 = n_dyn, c1 = n_dyn;
            if (n_dyn) {
                c0 = atomicAdd(&p.ctl->tile_ctr, 1u);
                c1 = atomicAdd(&p.ctl->tile_ctr, 1u);
            }
            auto admissible = [&](uint32_t tile) {
                uint32_t adm = 0xFFFFFFFFu;
                if (tile >= T) return 0u;
                if (p.mask) adm &= p.mask[tile];
                if (p.live) adm &= p.live[tile];
                return adm;
            };
            auto issue = [&](uint32_t tile) {
                const uint32_t row0 = tile * kI8TileRows;
                const uint32_t bytes = min(uint32_t(kI8TileRows), p.n - row0) * p.rec_bytes;
                mbar_arrive_expect_tx(&hdr->full[s], bytes);
                bulk_g2s(smem + p.stage_off + size_t(s) * p.stage_bytes, p.x8 + size_t(row0) * p.rec_bytes, bytes,
                         &hdr->full[s], pol);
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            };
            for (uint32_t it = 0; it < n_static; it++) {
                mbar_wait(&hdr->empty[s], ph ^ 1u);
                issue(blockIdx.x + it * G);
            }
            uint32_t a0[kI8DynChunk], a1[kI8DynChunk];
            if (p.dep_inputs) pdl_wait();
#pragma unroll
            for (uint32_t t = 0; t < kI8DynChunk; t++) a0[t] = (c0 < n_dyn) ? admissible(p.dyn_tile0 + c0 * kI8DynChunk + t) : 0u;
            while (c0 < n_dyn) {
                const uint32_t c2 = (c1 < n_dyn) ? atomicAdd(&p.ctl->tile_ctr, 1u) : n_dyn;
#pragma unroll
                for (uint32_t t = 0; t < kI8DynChunk; t++) a1[t] = (c1 < n_dyn) ? admissible(p.dyn_tile0 + c1 * kI8DynChunk + t) : 0u;
#pragma unroll
                for (uint32_t t = 0; t < kI8DynChunk; t++) {
                    const uint32_t tile = p.dyn_tile0 + c0 * kI8DynChunk + t;
                    if (tile >= T) break;
                    mbar_wait(&hdr->empty[s], ph ^ 1u);
                    hdr->tile_of[s] = tile;
                    hdr->adm_of[s] = a0[t];
                    issue(tile);
                }
                c0 = c1;
                c1 = c2;
#pragma unroll
                for (uint32_t t = 0; t < kI8DynChunk; t++) a0[t] = a1[t];
            }
            asm volatile("" ::"r"(c1));
            for (int i = 0; i < ncw; i++) {   // one stop marker per consumer warp (warp w owns the stages s % ncw == w)
                mbar_wait(&hdr->empty[s], ph ^ 1u);
                hdr->tile_of[s] = kNoTile;
                mbar_arrive(&hdr->full[s]);
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        return;
    }

    // ---- the normalised fp32 query goes to shared memory once (the exact re-scoring reads it) ----
    float4* qs = reinterpret_cast<float4*>(smem + p.q_off);
    if (p.dep_inputs) pdl_wait();
    {
        bool scaled;
        const float inv = i8_query_scale(p.q, p.d, p.ld4, p.normalize_q, lane, &scaled);   // same value in every warp
        for (int c = cw * kWarp + lane; c < p.ld4; c += ncw * kWarp) {
            float4 v = i8_load_q4(p.q, p.d, c);
            if (scaled) { v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
            qs[c] = v;
        }
    }
    named_bar_sync(1, ncw * 32);
    // ---- query: two int8 planes, their scales, ||q||, rq = ||q - q^||  (identical in every warp) ----
    const int J = (p.ld8 / 16 + 31) / 32;
    int4 q1[kI8MaxJ], q2[kI8MaxJ];
    float s1, s2, qnorm, rq;
    {
        float f[kI8MaxJ][16];
        float mx = 0.f, nr = 0.f;
#pragma unroll
        for (int j = 0; j < kI8MaxJ; j++) {
            const int c = lane + 32 * j;   // 16-element chunk
#pragma unroll
            for (int t = 0; t < 4; t++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < J && 4 * c + t < p.ld4) v = qs[4 * c + t];
                f[j][4 * t + 0] = v.x; f[j][4 * t + 1] = v.y; f[j][4 * t + 2] = v.z; f[j][4 * t + 3] = v.w;
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                nr = dot4(v, v, nr);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        nr = warp_allsum(nr);
        qnorm = sqrtf(nr) * 1.00001f;
        s1 = mx * (1.0f / 127.0f);
        const float inv1 = s1 > 0.f ? 1.0f / s1 : 0.f;
        float mx2 = 0.f;
        int qa[kI8MaxJ][16];
#pragma unroll
        for (int j = 0; j < kI8MaxJ; j++)
#pragma unroll
            for (int t = 0; t < 16; t++) {
                int qv = max(-127, min(127, __float2int_rn(f[j][t] * inv1)));
                qa[j][t] = qv;
                f[j][t] = fmaf(-s1, float(qv), f[j][t]);   // residual of plane 1
                mx2 = fmaxf(mx2, fabsf(f[j][t]));
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx2 = fmaxf(mx2, __shfl_xor_sync(0xFFFFFFFFu, mx2, o));
        s2 = mx2 * (1.0f / 127.0f);
        const float inv2 = s2 > 0.f ? 1.0f / s2 : 0.f;
        float r2 = 0.f;
#pragma unroll
        for (int j = 0; j < kI8MaxJ; j++) {
            uint32_t w1[4] = {0, 0, 0, 0}, w2[4] = {0, 0, 0, 0};
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const int qb = max(-127, min(127, __float2int_rn(f[j][t] * inv2)));
                const float res = fmaf(-s2, float(qb), f[j][t]);
                r2 = fmaf(res, res, r2);
                w1[t >> 2] |= uint32_t(uint8_t(int8_t(qa[j][t]))) << (8 * (t & 3));
                w2[t >> 2] |= uint32_t(uint8_t(int8_t(qb))) << (8 * (t & 3));
            }
            q1[j] = make_int4(int(w1[0]), int(w1[1]), int(w1[2]), int(w1[3]));
            q2[j] = make_int4(int(w2[0]), int(w2[1]), int(w2[2]), int(w2[3]));
        }
        r2 = warp_allsum(r2);
        rq = sqrtf(r2) * 1.00001f + 1e-30f;
    }
    // B_r = qhat_norm * e_r + bconst;  eps covers the fp32 rounding of the scan's own score (gamma_d ||q|| ||x||),
    // of `approx` below and of this bound's own arithmetic
    const float qhat_norm = (qnorm + rq) * 1.0001f;
    const float bconst = (rq * p.max_norm + (float(p.d) * 1.2e-7f + 8e-6f) * qnorm * p.max_norm) * 1.0001f;

    const int my_row = tile_row_of_lane(lane);
    const bool leader = (lane & 3) == 0;
    // Slot of p.best this warp reports to.  k <= 32: one slot per CTA (the warps of a CTA share it through
    // atomicMax; 148 slots make the threshold cheap to recompute and the k-th largest of 148 CTA-bests is
    // still about the k-th best row).  Larger k: one slot per warp (more distinct rows at the top).
    const bool slot_per_cta = p.k <= 32;
    const uint32_t gw = slot_per_cta ? blockIdx.x : blockIdx.x * uint32_t(ncw) + uint32_t(cw);
    uint32_t n_cand = 0;
    uint32_t my_best = 0u, published = 0u, thr = 0u;
    uint32_t done_tiles = 0, next_refresh = 1, seen_refreshes = 0;
    const uint32_t n_static = min(iters, p.static_iters);
    for (uint32_t it = cw;; it += ncw) {
        const int s = it % S;
        uint32_t tile, adm = 0xFFFFFFFFu;
        if (it < n_static) {
            tile = blockIdx.x + it * G;
            if (p.mask) adm &= p.mask[tile];   // issued before the data wait
            if (p.live) adm &= p.live[tile];
            mbar_wait(&hdr->full[s], (it / S) & 1u);
        } else {
            mbar_wait(&hdr->full[s], (it / S) & 1u);
            tile = hdr->tile_of[s];
            if (tile == kNoTile) break;
            adm = hdr->adm_of[s];
        }
        const uint32_t row0 = tile * kI8TileRows;
        const uint8_t* st = smem + p.stage_off + size_t(s) * p.stage_bytes;
        uint32_t upv[4];   // leaders: upper bound of "their" row of each 8-row group (0 = not admissible)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            int a1[8], a2[8];
#pragma unroll
            for (int r = 0; r < 8; r++) a1[r] = a2[r] = 0;
#pragma unroll
            for (int j = 0; j < kI8MaxJ; j++) {
                const int c = lane + 32 * j;
                if (j < J && c * 16 < p.ld8) {
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        const int4 v = *reinterpret_cast<const int4*>(st + size_t(g * 8 + r) * p.rec_bytes + c * 16);
                        a1[r] = __dp4a(v.x, q1[j].x, a1[r]); a1[r] = __dp4a(v.y, q1[j].y, a1[r]);
                        a1[r] = __dp4a(v.z, q1[j].z, a1[r]); a1[r] = __dp4a(v.w, q1[j].w, a1[r]);
                        a2[r] = __dp4a(v.x, q2[j].x, a2[r]); a2[r] = __dp4a(v.y, q2[j].y, a2[r]);
                        a2[r] = __dp4a(v.z, q2[j].z, a2[r]); a2[r] = __dp4a(v.w, q2[j].w, a2[r]);
                    }
                }
            }
            // the two planes are combined per lane in fp32 (each partial dot is far below 2^24, so the
            // conversions are exact; the few roundings of the combination are inside eps) and reduced ONCE
            float fa[8];
#pragma unroll
            for (int r = 0; r < 8; r++) fa[r] = fmaf(s1, float(a1[r]), s2 * float(a2[r]));
            const float dq = reduce8(fa, lane);
            const int rt = g * 8 + my_row;   // row of the tile this quad is responsible for
            uint32_t up = 0u;
            if (leader && row0 + uint32_t(rt) < p.n && ((adm >> rt) & 1u)) {
                const float2 m = *reinterpret_cast<const float2*>(st + size_t(rt) * p.rec_bytes + p.ld8);
                const float approx = m.x * dq;
                const float B = fmaf(qhat_norm, m.y, bconst);
                if (approx == approx) {
                    my_best = max(my_best, score_to_ord(approx - B));
                    up = max(score_to_ord(approx + B), 1u);
                } else {
                    up = 0xFFFFFFFFu;   // not a number: let the exact re-scoring decide (the fp32 scan drops NaN scores)
                }
            }
            upv[g] = up;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&hdr->empty[s]);   // the stage is free: everything below works from registers
        // publish this warp's best lower bound (monotone; one word per warp, so a reader never sees a torn value)
        const uint32_t wb = __reduce_max_sync(0xFFFFFFFFu, my_best);
        if (wb > published) {
            published = wb;
            if (lane == 0) {
                if (slot_per_cta) atomicMax(p.best + gw, wb);
                else *reinterpret_cast<volatile unsigned int*>(p.best + gw) = wb;
            }
        }
        my_best = wb;
        // Threshold = k-th largest of all warps' bests.  Recomputed after 1, 2, 4, ... tiles of this warp, then
        // every 32 -- unless another warp of the CTA has done it since this warp last looked (the result is
        // shared through hdr->thr).  After the FIRST tile the other warps are only just publishing: give them a
        // moment, else every warp's first 32 rows would all become candidates.
        ++done_tiles;
        const unsigned int cta_refreshes = *reinterpret_cast<volatile unsigned int*>(&hdr->refreshes);
        if (done_tiles >= next_refresh) {
            next_refresh = done_tiles < 32 ? done_tiles * 2 : done_tiles + 32;
            if (done_tiles == 1 || cta_refreshes == seen_refreshes) {
                uint32_t t = i8_threshold(p.best, p.nbest, p.k, lane);
                if (done_tiles == 1)
                    for (int spin = 0; t == 0u && spin < 6; spin++) {
                        __nanosleep(500);
                        t = i8_threshold(p.best, p.nbest, p.k, lane);
                    }
                if (lane == 0) {
                    atomicMax(&hdr->thr, t);
                    atomicAdd(&hdr->refreshes, 1u);
                }
                thr = max(thr, t);
            }
            seen_refreshes = *reinterpret_cast<volatile unsigned int*>(&hdr->refreshes);
        }
        thr = max(thr, *reinterpret_cast<volatile unsigned int*>(&hdr->thr));
        // Rows the int8 pass cannot rule out (a few hundred per search): re-score them right here from the fp32
        // master row -- per-lane chunk order + xor butterfly == the fp32 scan's reduce8 tree, so the score is
        // bit-identical to the scan's -- and keep the exact key if it reaches the threshold.
#pragma unroll 1
        for (int g = 0; g < 4; g++) {
            unsigned m = __ballot_sync(0xFFFFFFFFu, upv[g] != 0u && upv[g] >= thr);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t row = row0 + uint32_t(g * 8 + tile_row_of_lane(src));
                const float4* xr = reinterpret_cast<const float4*>(p.x) + size_t(row) * p.ld4;
                float acc = 0.f;
                for (int c = lane; c < p.ld4; c += kWarp) acc = dot4(ldg_stream(xr + c), qs[c], acc);
                acc = warp_allsum(acc);
                n_cand++;
                if (lane == 0 && acc == acc && score_to_ord(acc) >= thr) {
                    const unsigned pos = atomicAdd(&p.ctl->surv_cnt, 1u);
                    if (pos < kI8SurvCap) p.surv[pos] = make_key(acc, row);
                }
            }
        }
    }
    // ---- tail: the last CTA to finish sorts the survivors and writes (D, I) ----
    if (lane == 0 && n_cand) atomicAdd(&p.ctl->cand_cnt, n_cand);
    __threadfence();
    named_bar_sync(1, ncw * 32);
    if (cw == 0 && lane == 0) {
        const bool last = atomicAdd(&p.ctl->ticket, 1u) == G - 1;
        hdr->last_flag = last;
        if (last) {
            // ONE thread reads the counts and leaves the shared state clean for the next search on this workspace
            // (every other CTA is done with it); the rest of the CTA takes the counts from shared memory after the
            // barrier -- a thread reading the global counters itself could already see them reset
            __threadfence();
            const unsigned int ns_now = *reinterpret_cast<volatile unsigned int*>(&p.ctl->surv_cnt);
            const unsigned int nc_now = *reinterpret_cast<volatile unsigned int*>(&p.ctl->cand_cnt);
            hdr->tail_ns = ns_now;
            hdr->tail_nc = nc_now;
            p.ctl->last_cand = nc_now;
            p.ctl->last_surv = ns_now;
            p.ctl->cand_cnt = 0u;
            p.ctl->surv_cnt = 0u;
            p.ctl->ticket = 0u;
            p.ctl->tile_ctr = 0u;
        }
    }
    named_bar_sync(1, ncw * 32);
    if (!hdr->last_flag) return;
    __threadfence();
    const int tid = cw * kWarp + lane, nthr = ncw * kWarp;
    const unsigned int ns = hdr->tail_ns;
    for (uint32_t i = tid; i < p.nbest; i += nthr) p.best[i] = 0u;
    if (ns > kI8SurvCap) {
        if (tid == 0) {   // the fp32 scan answers this query instead (conditional launch behind us, or the host re-runs it)
            p.ctl->overflow = 1u;
            if (p.ovf_host) {
                *reinterpret_cast<volatile unsigned int*>(p.ovf_host) = 1u;
                __threadfence_system();
            }
        }
        return;
    }
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem + p.stage_off);   // the ring is idle now
    for (uint32_t i = tid; i < ns; i += nthr) sk[i] = __ldcg(p.surv + i);
    named_bar_sync(1, nthr);
    // the k best of the survivors: register sort of a short list, block-wide k-th-element selection of a long one
    int cnt = 0;
    uint64_t* fin = block_select_topk(sk, ns, p.k, hdr->cnts4, tid, nthr, cw, lane, &cnt);
    if (!fin) return;
    if (p.xchg) {
        // sharded search: same protocol as the fp32 scan's tail (scan.cuh finish_scan) -- send, publish, wait, merge
        xchg_send(p.xchg, p.xchg_seq, 0, fin, cnt, p.k, lane);
        xchg_publish_and_wait(p.xchg, p.xchg_seq, lane);
        xchg_merge(p.xchg, p.xchg_seq, 0, fin == sk ? sk + kSelectMax : sk, select_cap(p.k), p.k, p.outD, p.outI, lane);
        return;
    }
    write_results(fin, cnt, p.k, p.outD, p.outI, p.label_offset, lane);
}

}  // namespace mvdb
